"""The reference's ONE test (ref:tests/test_prepare_albedo_alpha.py:24-71), run against this repository's prepare stage with the same inputs and the
same assertion: an 8-bit normal PNG + a 16-bit albedo PNG, no mask file, scaling_mode="cameras" -> the prepared albedo's alpha channel must be fully
opaque AT ITS OWN BIT DEPTH (the bug it guards: a 0/255 mask derived from the 8-bit normal pasted onto the 16-bit albedo, alpha 255/65535)."""
import importlib
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


class _Log:
    def info(self, m): pass
    def warning(self, m): pass


def _scene(tmp, normal_dtype, albedo_dtype):
    h, w = 16, 16
    normal = np.full((h, w, 3), 128 if normal_dtype == np.uint8 else 32896, dtype=normal_dtype)
    normal_path = os.path.join(tmp, "n0.png"); cv2.imwrite(normal_path, normal)
    albedo = np.full((h, w, 3), 30000 if albedo_dtype == np.uint16 else 117, dtype=albedo_dtype)
    albedo_path = os.path.join(tmp, "a0.png"); cv2.imwrite(albedo_path, albedo)
    K = np.eye(3, dtype=np.float32)
    c2w = np.eye(4, dtype=np.float32); c2w[2, 3] = 3.0
    return {"views": [{"c2w": c2w, "K": K, "normal_path": normal_path, "albedo_path": albedo_path, "mask_path": None}],      # mask None -> full mask
            "landmarks": None, "image_width": w, "image_height": h}


@pytest.mark.parametrize("normal_dtype,albedo_dtype", [(np.uint8, np.uint16), (np.uint16, np.uint8), (np.uint8, np.uint8), (np.uint16, np.uint16)])
def test_prepared_albedo_and_normal_alpha_are_opaque_at_their_own_bit_depth(pkg, tmp_path, normal_dtype, albedo_dtype):
    prepare = importlib.import_module("rnb_neus2_b200.prepare")
    data = _scene(str(tmp_path), normal_dtype, albedo_dtype)
    out = os.path.join(str(tmp_path), "prepared")
    prepare.prepare_testbed_data(data, out, _Log(), scaling_mode="cameras")
    for sub in ("albedos", "normals"):
        prep = cv2.imread(os.path.join(out, sub, "00000.png"), cv2.IMREAD_UNCHANGED)
        assert prep is not None and prep.ndim == 3 and prep.shape[2] == 4, sub
        expected = 65535 if prep.dtype == np.uint16 else 255
        assert int(prep[:, :, 3].max()) == expected and int(prep[:, :, 3].min()) == expected, (sub, prep.dtype)
    # and the library's own PNG decoder (what training reads) sees a full-scale mask either way
    for sub in ("albedos", "normals"):
        px = pkg.load_png_rgba16(os.path.join(out, sub, "00000.png"))
        assert px.shape == (16, 16, 4) and set(np.unique(px[:, :, 3]).tolist()) == {65535}


def test_reference_module_agrees_on_the_same_case(tmp_path):
    """where the reference checkout is present (this container, not the GPU box): its own prepare on the same inputs writes the same files"""
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "rnb_neus2")):
        pytest.skip("reference checkout not present")
    import hashlib
    import sys
    import rnb_loader
    rnb_loader.load_package()
    ours = importlib.import_module("rnb_neus2_b200.prepare")
    sys.path.insert(0, ref_root)
    try:
        theirs = importlib.import_module("rnb_neus2.prepare")
    finally:
        sys.path.remove(ref_root)
    data = _scene(str(tmp_path), np.uint8, np.uint16)
    digests = []
    for name, mod in (("ours", ours), ("ref", theirs)):
        out = os.path.join(str(tmp_path), name)
        mod.prepare_testbed_data(data, out, _Log(), scaling_mode="cameras")
        files = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(out) for f in fs)
        digests.append([(os.path.relpath(f, out), hashlib.sha256(open(f, "rb").read()).hexdigest()) for f in files])
    assert digests[0] == digests[1]
