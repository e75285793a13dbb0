"""The reference pipeline's stage 2, end to end on this repo's API (what `./build/testbed --scene DIR --maxiter N --save-snapshot
--save-mesh --resolution R` does, src/main.cu:300-470): scene directory -> training -> snapshot -> mesh file; then the stage hand-off
of rnb_neus2/pipeline.py: a second Testbed resumes from the snapshot and extracts the same mesh."""
import numpy as np
import pytest
import ref_scene
from common import FULL, product_config

pytestmark = pytest.mark.gpu


def test_scene_dir_to_snapshot_and_mesh(pkg, scene_mod, tmp_path):
    views = scene_mod.make_scene(12, 128, 128, with_albedo=True)
    n2w = np.eye(4); n2w[:3, :3] *= 2.0; n2w[:3, 3] = (1.0, 2.0, 3.0)
    ref_scene.write_scene(str(tmp_path / "scene"), views, n2w=n2w)
    mk = lambda: pkg.Testbed(product_config(pkg, FULL, rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    t = mk(); t.init_params()
    meta = t.load_training_data_dir(str(tmp_path / "scene"))
    losses = [t.train().loss for _ in range(120)]
    assert np.isfinite(losses).all() and np.mean(losses[-10:]) < np.mean(losses[:5])
    snap_path = tmp_path / "snapshot_120.msgpack"; obj = tmp_path / "mesh_120.obj"
    t.save_snapshot(snap_path, {"encoding": {"n_levels": 14}})
    info = t.compute_and_save_marching_cubes_mesh(obj, 128, nerf_scale=meta["scale"], nerf_offset=meta["offset"], n2w_s=meta["n2w_s"], n2w_t=meta["n2w_t"], from_na=meta["from_na"])
    assert info["n_verts"] > 1000 and info["n_indices"] > 3000
    lines = open(obj).read().split("\n")
    v = np.array([[float(x) for x in l.split()[1:4]] for l in lines if l.startswith("v ")])
    f = np.array([[int(x.split("/")[0]) for x in l.split()[1:]] for l in lines if l.startswith("f ")])
    assert v.shape[0] == info["n_verts_padded"] and f.shape[0] == info["n_indices"] // 3 and f.min() == 1 and f.max() == info["n_verts"]
    # world frame: n2w_s * ((p - offset) / scale) + n2w_t of positions inside the unit cube -> inside [-2, 2]^3 + t
    real = v[:info["n_verts"]]
    assert np.all(real >= np.array([1.0, 2.0, 3.0]) - 2.0 - 1e-4) and np.all(real <= np.array([1.0, 2.0, 3.0]) + 2.0 + 1e-4)
    # stage hand-off: a fresh context resumes from the snapshot and gets the same surface
    u = mk(); u.load_training_data_dir(str(tmp_path / "scene")); u.load_snapshot(snap_path)
    info2 = u.marching_cubes(128, use_ema=True)
    assert info2["n_verts"] == info["n_verts"] and info2["n_indices"] == info["n_indices"]
    assert np.array_equal(u.mesh_download()["V"], t.mesh_download()["V"])
    assert np.isfinite(u.train().loss)
