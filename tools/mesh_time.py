#!/usr/bin/env python
"""Time the mesh path (SDF sweep -> marching cubes -> normals -> colours -> OBJ text) on the default network.  GPU box only."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader
from common import FULL, product_config
import torch

pkg = rnb_loader.load_package(); scene = rnb_loader.load_scene()
views = scene.make_scene(12, 256, 256, with_albedo=True)
t = pkg.Testbed(product_config(pkg, FULL, rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
t.init_params(); t.load_training_data(views)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 300):
    t.train(want_stats=False)
rows = []
for res in (256, 512, 1024):
    n = res ** 3
    sd = torch.empty(n, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t.sdf_on_grid_device((res,) * 3, (0, 0, 0), (1, 1, 1), sd.data_ptr(), use_ema=True); torch.cuda.synchronize()      # warm
    ev[0].record(); t.sdf_on_grid_device((res,) * 3, (0, 0, 0), (1, 1, 1), sd.data_ptr(), use_ema=True); ev[1].record(); torch.cuda.synchronize()
    w0 = time.time(); info = t.marching_cubes_from_density(sd.data_ptr(), (res,) * 3, with_colors=False); torch.cuda.synchronize(); t_mc = time.time() - w0
    w0 = time.time(); info = t.marching_cubes_from_density(sd.data_ptr(), (res,) * 3, with_colors=True); torch.cuda.synchronize(); t_mcc = time.time() - w0
    row = {"res": res, "sdf_sweep_ms": round(ev[0].elapsed_time(ev[1]), 3), "extract_normals_ms": round(t_mc * 1e3, 3), "extract_normals_colors_ms": round(t_mcc * 1e3, 3), **info}
    row["roofline"] = {"alg_bytes_per_point": 4, "GBps": round(4 * n / ((info["stage_ms"]["bits_count_scan"] + info["stage_ms"]["vertices_normals_faces"]) * 1e-3) / 1e9, 1)}
    if res <= 512:
        path = "/tmp/mesh_%d.obj" % res
        w0 = time.time(); nb = t.save_mesh(path, 0.5, (0.5, 0.5, 0.5), 1.0, (0, 0, 0), True); row["obj_write_ms"] = round((time.time() - w0) * 1e3, 3); row["obj_bytes"] = nb
        for key in ("pipeline_first_ms", "pipeline_ms"):        # first call at a resolution grows the scratch buffers
            w0 = time.time(); t.compute_and_save_marching_cubes_mesh(path, res, nerf_scale=0.5, nerf_offset=(0.5, 0.5, 0.5), from_na=True); row[key] = round((time.time() - w0) * 1e3, 3)
        os.remove(path)
    del sd
    rows.append(row); print(json.dumps(row)); sys.stdout.flush()
json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mesh_time.json", "w"), indent=1)
