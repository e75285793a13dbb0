"""The prepare stage (scene normalisation + testbed folder layout) against the reference's own modules: tests/golden/ref_prepare.npz
was produced by rnb_neus2/prepare.py + scaling.py (tests/golden/make_prepare_golden.py) on the same inputs, for every scaling mode.  Compared
here: the transform.json text (character for character), every PNG written (sha256), the log lines, the returned scaling — and that the result
is a scene directory this repository's loader reads."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_prepare_golden import MODES, run  # noqa: E402

GOLD = json.loads(bytes(np.load(os.path.join(ROOT, "tests", "golden", "ref_prepare.npz"))["gold"]).decode())


@pytest.mark.parametrize("mode", MODES)
def test_prepare_matches_the_reference_module(pkg, tmp_path, mode):
    mod = importlib.import_module("rnb_neus2_b200.prepare")
    got = run(mod.prepare_testbed_data, str(tmp_path), mode)
    want = GOLD[mode]
    assert got["n_frames"] == want["n_frames"] == 7                      # view 6 has no normal map and is skipped
    assert got["factor"] == want["factor"] and got["center"] == want["center"]
    assert got["matrix"] == want["matrix"] and got["n2w"] == want["n2w"]
    assert got["transform"] == want["transform"], "transform.json differs from the reference's"
    assert got["files"] == want["files"], "PNG files are not byte-identical to the reference's"
    assert got["log"] == want["log"]


def test_prepared_scene_is_what_the_loader_reads(pkg, tmp_path):
    """prepare -> dataset.load_transforms -> rnb_load_png_rgba16: the folder is a valid scene for this repository's ingest (and carries
    the mask in the alpha channel at full scale for both 8- and 16-bit sources)"""
    from albedo_scene import write_prepare_inputs
    mod = importlib.import_module("rnb_neus2_b200.prepare")
    ds = importlib.import_module("rnb_neus2_b200.dataset")

    class Log:
        def info(self, m): pass
        def warning(self, m): pass
    data = write_prepare_inputs(str(tmp_path))
    out = tmp_path / "scene"
    ret = mod.prepare_testbed_data(data, str(out), Log(), scaling_mode="auto")
    tj = json.loads((out / "transform.json").read_text())
    assert tj["from_na"] is True and len(tj["frames"]) == ret["n_frames"] == 7
    assert np.allclose(np.array(tj["n2w"]) @ np.array(ret["scale_matrix"], np.float64), np.eye(4), atol=1e-5)
    # every camera centre lies inside the normalised volume the training loop samples
    centers = np.array([np.array(f["transform_matrix"])[:3, 3] for f in tj["frames"]])
    assert np.all(np.linalg.norm(centers, axis=1) < 4.0)
    meta = ds.load_transforms(str(out))
    assert len(meta["views"]) == 7
    for sub in ("normals", "albedos"):
        for name in ("00001.png", "00004.png"):                          # 8-bit and 16-bit albedo sources
            px = pkg.load_png_rgba16(str(out / sub / name))
            assert px.shape[2] == 4 and set(np.unique(px[:, :, 3]).tolist()) <= {0, 65535}
    with pytest.raises(RuntimeError):
        mod.prepare_testbed_data({"views": [], "landmarks": None, "image_width": 4, "image_height": 4}, str(tmp_path / "e"), Log(), scaling_mode="pcd")
