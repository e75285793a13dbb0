#!/usr/bin/env python
"""Distribution of marched / kept samples per ray on the bench workload (profiling aid)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader, bench
pkg = rnb_loader.load_package()
views, _ = bench.build_views(96, 1600, 1200, False)
t = pkg.Testbed(pkg.default_config(rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=1))
t.init_params(); t.load_training_data(views)
out = {}
for target in (100, 300, 700, 1500):
    while t.get_train_state()[0] < target:
        t.train(want_stats=False)
    s = t.train()
    m, k = t.ray_counts()
    m = m.astype(np.int64); k = k.astype(np.int64)
    row = {"step": target, "rays_kept": int(m.size), "marched": int(m.sum()), "kept": int(k.sum()), "marched_max": int(m.max()), "kept_max": int(k.max()),
           "kept_pct": [int(np.percentile(k, q)) for q in (50, 90, 99)], "marched_pct": [int(np.percentile(m, q)) for q in (50, 90, 99)]}
    # two-phase pass A (evaluate the first K1 samples of every ray, then the rest only for rays whose transmittance is still above the
    # cut): samples to evaluate, exactly and at warp (32 consecutive sample slots, ray-major order) / tile (128) granularity
    base = np.concatenate([[0], np.cumsum(m)]); total = int(base[-1])
    ray_of = np.repeat(np.arange(m.size), m); j = np.arange(total) - base[ray_of]
    for K1 in (32, 64, 96, 128, 192):
        alive = k > K1            # still alive after the first K1 samples
        need = (j < K1) | alive[ray_of]
        row["work_K1_%d" % K1] = int(need.sum())
        for g in (32, 128):
            pad = (-total) % g
            blocks = np.concatenate([need, np.zeros(pad, bool)]).reshape(-1, g).any(axis=1)
            row["work_K1_%d_gran%d" % (K1, g)] = int(blocks.sum() * g)
    row["ideal"] = int(np.minimum(k + 1, m).sum())      # samples up to and including the one that crosses the cut
    out[target] = row
    print(json.dumps(row))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ray_stats.json"), "w"), indent=1)
