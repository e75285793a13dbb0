#!/usr/bin/env python
"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck), run on the B200 box by tools/gpu_sanitize.sh: the default-width network's
forward and backward (tcgen05 kernels: mbarriers, TMEM reuse across tiles, warp-specialised scatter hand-off) on a ray-ordered batch, and a few full
training steps with occupancy refresh, adaptive controller and mesh extraction on a tiny scene.  EVIDENCE TOOLING."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader
pkg = rnb_loader.load_package(); scene = rnb_loader.load_scene()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
MID = dict(n_levels=14, log2_hashmap_size=15, base_resolution=16, top_resolution=2048.0, sdf_n_neurons=64, sdf_n_hidden_layers=1, rgb_n_neurons=64, rgb_n_hidden_layers=2)
rs = np.random.RandomState(0)
if what in ("all", "network"):
    t = pkg.Testbed(pkg.default_config(**MID))
    t.init_params()
    t.set_train_state(700, 512)
    n = 3 * 128 + 37                                   # several tiles per CTA are not reached at this size on 148 SMs: ragged last tile
    coords = rs.rand(n, 7).astype(np.float32)
    out, nrm = t.stage_forward(coords)
    dout = (rs.randn(n, 16) * 0.05).astype(np.float16).astype(np.float32); dout[:, 11:] = 0
    g = t.stage_backward(coords, dout, n - 50)
    print("network: forward", out.shape, "finite", bool(np.isfinite(out).all()), "grad norm", float(np.linalg.norm(g)))
    n = 148 * 128 * 2 + 5                              # two tiles per CTA: the mailbox / barrier phases wrap
    coords = rs.rand(n, 7).astype(np.float32)
    dout = (rs.randn(n, 16) * 0.05).astype(np.float16).astype(np.float32); dout[:, 11:] = 0
    g = t.stage_backward(coords, dout, n)
    print("network: two tiles per CTA, grad norm", float(np.linalg.norm(g)))
    t.close()
if what in ("all", "train"):
    views = scene.make_scene(4, 64, 64, with_albedo=True)
    t = pkg.Testbed(pkg.default_config(rays_per_batch=256, pin_rays_per_batch=0, target_batch_size=1 << 14, **MID), pkg.default_flags(no_albedo=0, light_mode=-2))
    t.init_params(); t.load_training_data(views)
    for _ in range(6):
        st = t.train()
    print("train: step", st.training_step, "loss", st.loss, "rays next", st.rays_per_batch_next)
    for _ in range(3):
        t.train(want_stats=False)
    info = t.marching_cubes((32, 32, 32))
    print("mesh:", info)
    t.close()
print("sanitize_case done")
