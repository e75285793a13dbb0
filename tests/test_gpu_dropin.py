"""Drop-in boundary on hardware: the reference's own CLI (src/main.cu, unmodified) linked with the six hooks of shim/rnb_testbed_shim.h against
librnb_b200.so (`oracle/_ref/bin/testbed_rnb`, built by `make -C oracle -f Makefile.ref shim`) trains, writes a snapshot and a mesh with
the reference's argv (ref:rnb_neus2/pipeline.py:27-53, ref:src/main.cu:283-469), and resumes from its own snapshot like stage 2 does."""
import glob
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "testbed_rnb")


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/bin/testbed_rnb not built (make -C oracle -f Makefile.ref shim)")
def test_testbed_rnb_trains_saves_and_resumes(pkg, scene_mod, tmp_path):
    import ref_scene
    snap = importlib.import_module(pkg.__name__ + ".snapshot")
    views = scene_mod.make_scene(8, 128, 96, with_albedo=False)
    sd = str(tmp_path / "scene")
    ref_scene.write_scene(sd, views)
    cmd = [BIN, "--scene", sd + "/", "--maxiter", "200", "--no-gui", "--mask-weight", "1.0", "--save-snapshot", "--no-albedo"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    its = [(int(a), float(b)) for a, b in re.findall(r"iteration=(\d+) loss=([0-9.eE+-]+)", r.stdout)]
    # src/main.cu:444-449 prints after every frame() that returns true, when the step is a multiple of 100; the frame that reaches --maxiter returns false
    assert [i for i, _ in its] == [100] and all(np.isfinite(v) and v > 0 for _, v in its)
    s1 = os.path.join(sd, "output", "snapshot_200.msgpack")
    cfg = snap.read_snapshot(s1)
    p = snap.parse_snapshot(cfg)
    assert p["training_step"] == 200 and p["params_fp16"].size == 10559396 and np.isfinite(p["params_fp16"].astype(np.float32)).all()
    assert p["density_grid"].size == 128 ** 3 and (p["density_grid"] > 0).any()
    assert 128 <= p["rays_per_batch"] <= (1 << 18) and p["rays_per_batch"] % 128 == 0          # the adaptive controller ran (testbed_nerf.cu:3554-3555)
    # stage 2: resume from the snapshot, different light basis, mesh + snapshot out
    cmd2 = [BIN, "--scene", sd + "/", "--maxiter", "350", "--no-gui", "--mask-weight", "1.0", "--opti-lights", "--snapshot", s1, "--resolution", "64", "--save-mesh", "--save-snapshot", "--no-albedo"]
    r2 = subprocess.run(cmd2, capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    its2 = [int(a) for a, _ in re.findall(r"iteration=(\d+) loss=([0-9.eE+-]+)", r2.stdout)]
    assert its2 == [300]                                                                       # resumed at 200, not at 0 (which would also print 100 and 200)
    p2 = snap.parse_snapshot(snap.read_snapshot(os.path.join(sd, "output", "snapshot_350.msgpack")))
    assert p2["training_step"] == 350
    assert not np.array_equal(p2["params_fp16"], p["params_fp16"])                             # it trained
    m = glob.glob(os.path.join(sd, "output", "mesh_350.obj"))
    assert m
    v = np.array([[float(x) for x in line.split()[1:4]] for line in open(m[0]) if line.startswith("v ")])
    v = v[np.abs(v).sum(1) > 0]
    assert len(v) > 100
    # the surface is the scene's ellipsoid (world frame = (ngp - 0.5) / 0.5): coarse check after 350 steps at a 64^3 lattice
    q = v / (np.asarray(scene_mod.AXES, np.float64) / 0.5)
    assert np.median(np.abs(np.linalg.norm(q, axis=1) - 1.0)) < 0.15
