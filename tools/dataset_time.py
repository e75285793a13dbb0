#!/usr/bin/env python
"""Time dataset ingest: this repo's loader vs the reference's load_nerf on the same scene directory.  GPU box only."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["RNB_DATASET_DEBUG"] = "1"
import rnb_loader, ref_scene
from common import FULL, product_config

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dataset_time.json"
n_views, w, h = 24, 800, 600
pkg = rnb_loader.load_package(); scene = rnb_loader.load_scene()
views = scene.make_scene(n_views, w, h, with_albedo=True)
sdir = "/tmp/dataset_time_scene"
t0 = time.time(); ref_scene.write_scene(sdir, views, workers=os.cpu_count() or 4); t_write = time.time() - t0
raw_bytes = n_views * 2 * w * h * 8
file_bytes = sum(os.path.getsize(os.path.join(sdir, d, f)) for d in ("normals", "albedos") for f in os.listdir(os.path.join(sdir, d)))
rows = {"views": n_views, "resolution": [w, h], "raw_MB": round(raw_bytes / 1e6, 1), "png_MB": round(file_bytes / 1e6, 1), "host_cores": os.cpu_count(), "scene_write_s": round(t_write, 2)}
t = pkg.Testbed(product_config(pkg, FULL, rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
t.init_params()
for key, th in (("ours_all_cores_s", 0), ("ours_all_cores_again_s", 0), ("ours_1_thread_s", 1)):
    t0 = time.time(); t.load_training_data_dir(sdir, threads=th); rows[key] = round(time.time() - t0, 4)
st = t.train(); rows["first_step_loss"] = float(st.loss)
harness = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")
dump = "/tmp/dataset_time_dump"; os.makedirs(dump, exist_ok=True)
cfg = os.path.join(ROOT, "oracle", "_ref", "configs", "nerf", "base.json")
r = subprocess.run([harness, sdir + "/", cfg, dump, "1", "--time-only", "--pin-rays", "256"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
if r.returncode == 0:
    rows["reference_load_training_data_s"] = float(ref_scene.read_meta(os.path.join(dump, "meta.txt"))["load_seconds"])
else:
    rows["reference_error"] = r.stdout[-500:]
rows["ours_MBps_raw"] = round(raw_bytes / 1e6 / rows["ours_all_cores_again_s"], 1)
print(json.dumps(rows)); json.dump(rows, open(out, "w"), indent=1)
