#!/usr/bin/env python
"""Throughput of the marching-cubes SDF sweep (rnb_sdf_on_grid) on one GPU."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, rnb_loader
pkg = rnb_loader.load_package()
t = pkg.Testbed(pkg.default_config())
t.init_params()
res_list = [(256,) * 3, (512,) * 3, (1024,) * 3]
out = torch.empty(1024 ** 3, device="cuda")
rows = []
for res in res_list:
    n = res[0] * res[1] * res[2]
    for it in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); t.sdf_on_grid_device(res, (0, 0, 0), (1, 1, 1), out.data_ptr(), use_ema=False); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    rows.append({"res": res[0], "points": n, "ms": round(ms, 3), "Gpoints_per_s": round(n / ms / 1e6, 2), "finite": bool(torch.isfinite(out[:n]).all())})
    print(json.dumps(rows[-1]))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sdf_grid_time.json"), "w"), indent=1)
