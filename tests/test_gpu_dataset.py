"""Dataset ingest on the GPU box: a scene directory read through rnb_load_dataset_images (PNG decode on host
threads into pinned staging + upload) trains exactly like the same views handed over in memory."""
import time
import numpy as np
import pytest
import ref_scene
from common import SMALL, product_config

pytestmark = pytest.mark.gpu


def test_scene_directory_trains_like_in_memory_views(pkg, scene_mod, tmp_path):
    views = scene_mod.make_scene(6, 160, 120, with_albedo=True)
    ref_scene.write_scene(str(tmp_path), views)
    mk = lambda: pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=512, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    a = mk(); a.init_params(); a.load_training_data(views)
    b = mk(); b.init_params()
    t0 = time.time(); meta = b.load_training_data_dir(str(tmp_path), threads=4); dt = time.time() - t0
    assert len(meta["views"]) == 6 and meta["from_na"] and dt < 30
    # first step: same parameters, same pixels, same cameras -> identical rays, samples and per-ray losses
    sa, sb = a.train(), b.train()
    assert (sa.n_samples, sa.n_samples_compacted) == (sb.n_samples, sb.n_samples_compacted)
    ra, la = a.ray_losses(); rb, lb = b.ray_losses()
    assert np.array_equal(ra, rb) and np.array_equal(la, lb)
    # later steps differ only by the order of the gradient atomics
    for _ in range(3):
        sa, sb = a.train(), b.train()
        assert sa.n_samples == sb.n_samples and abs(sa.loss - sb.loss) < 1e-3 * max(abs(sa.loss), 1e-6)
    with pytest.raises(pkg.RnbError, match="image not found"):
        import json, os
        j = json.load(open(tmp_path / "transform.json")); j["frames"][2]["normal_path"] = "normals/nope.png"
        os.makedirs(tmp_path / "bad", exist_ok=True); json.dump(j, open(tmp_path / "bad" / "transform.json", "w"))
        mk().load_training_data_dir(str(tmp_path / "bad"))
