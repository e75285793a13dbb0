#!/bin/bash
# Drop-in run of the reference's own CLI with the training path routed through librnb_b200.so (oracle/_ref/bin/testbed_rnb, built by
# `make -C oracle -f Makefile.ref shim`) next to the stock binary on the same scene directory: same argv as run_pipeline.py's stage 1
# (rnb_neus2/pipeline.py:27-53).  NOT YET RUN ON HARDWARE (written after the round-1 GPU budget was spent): first thing to run in round 2.
#   gpurun --timeout 600 -- 'bash tools/gpu_shim_run.sh shim1 2000 256'
TAG=${1:-shim}; ITERS=${2:-2000}; RES=${3:-256}
O=gpurun_out/$TAG; mkdir -p $O
python - <<PY
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rnb_loader, ref_scene
scene = rnb_loader.load_scene()
views = scene.make_scene(24, 400, 300, with_albedo=False)
for name in ("stock", "rnb"):
    ref_scene.write_scene(os.path.join("$O", "scene_" + name), views, workers=8)
PY
for name in stock rnb; do
  bin=oracle/_ref/bin/testbed; [ $name = rnb ] && bin=oracle/_ref/bin/testbed_rnb
  /usr/bin/time -f "$name wall %e s" timeout 500 $bin --scene $O/scene_$name --maxiter $ITERS --no-gui --no-albedo --save-mesh --save-snapshot --resolution $RES > $O/$name.log 2>&1
  echo "$name rc=$?"; tail -3 $O/$name.log | cut -c1-200; ls -la $O/scene_$name/output | tail -3
done
python - <<PY
import glob, numpy as np
from scipy.spatial import cKDTree
def verts(p):
    return np.array([[float(x) for x in l.split()[1:4]] for l in open(p) if l.startswith("v ")])
a = verts(glob.glob("$O/scene_stock/output/mesh_*.obj")[0]); b = verts(glob.glob("$O/scene_rnb/output/mesh_*.obj")[0])
d_ab = cKDTree(b).query(a)[0]; d_ba = cKDTree(a).query(b)[0]
print("vertices stock %d rnb %d; nearest-vertex distance stock->rnb mean %.5f max %.5f, rnb->stock mean %.5f max %.5f (scene radius ~1)" % (len(a), len(b), d_ab.mean(), d_ab.max(), d_ba.mean(), d_ba.max()))
PY
