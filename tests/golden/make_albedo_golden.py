#!/usr/bin/env python
"""Generates tests/golden/ref_albedo_scaling.npz by running the REFERENCE's own module (rnb_neus2/albedo_scaling.py, imported from
/root/reference; run in the build container, the fixture travels).

The reference imports `trimesh`, which this image does not have.  Only its two ray queries are used (`trimesh.load_mesh(path)` and
`mesh.ray.intersects_location(ray_origins, ray_directions, multiple_hits)`), so a stand-in module with exactly that surface is put in
sys.modules, backed by brute force over all triangles in binary64 (oracle/orc_albedo.py).  Everything else — image and camera loading,
sampling with np.random.choice, projection, scipy interpolation, ratios, medians, chaining, normalisation, and scale_and_save_albedos —
is the reference's code, unmodified.  The fixture therefore pins the restatement and the product's host logic against the reference
for this stage, up to the choice of intersector.

Scene: tests/albedo_scene.py (8 views 96x80 around a sphere, per-view gains); files written as 16-bit PNGs + transform.json + OBJ.
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import orc_albedo  # noqa: E402
from albedo_scene import write_scene_files  # noqa: E402


def _read_obj(path):
    vs, fs = [], []
    for line in open(path):
        if line.startswith("v "):
            vs.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            fs.append([int(tok.split("/")[0]) - 1 for tok in line.split()[1:4]])
    return np.array(vs, np.float32), np.array(fs, np.uint32)


class _Ray:
    def __init__(self, verts, tris):
        self.v, self.f = verts, tris

    def intersects_location(self, ray_origins, ray_directions, multiple_hits=True):
        o = np.asarray(ray_origins, np.float64); d = np.asarray(ray_directions, np.float64)
        t = orc_albedo.ray_params_bruteforce(self.v, self.f, o, d)
        with np.errstate(invalid="ignore"):
            t = np.where(t > 0, t, np.nan)
        if multiple_hits:
            ri, ti = np.where(~np.isnan(t))
        else:
            has = ~np.all(np.isnan(t), axis=1)
            ri = np.where(has)[0]
            ti = np.nanargmin(np.where(has[:, None], t, 0.0), axis=1)[ri]
        return o[ri] + d[ri] * t[ri, ti][:, None], ri, ti


def install_trimesh_stand_in():
    m = types.ModuleType("trimesh")

    def load_mesh(path):
        v, f = _read_obj(path)
        return types.SimpleNamespace(vertices=v, faces=f, ray=_Ray(v, f))
    m.load_mesh = load_mesh
    sys.modules["trimesh"] = m


def main():
    if not os.path.isdir("/root/reference/rnb_neus2"):
        raise SystemExit("the reference is not present: run this in the build container")
    install_trimesh_stand_in()
    sys.path.insert(0, "/root/reference")
    from rnb_neus2 import albedo_scaling as ref
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        info = write_scene_files(tmp)
        for seed, ns in ((3, 400), (11, 150)):
            np.random.seed(seed)
            out["ratios_seed%d_n%d" % (seed, ns)] = ref.compute_albedo_scale_ratios(os.path.join(tmp, "albedos"), os.path.join(tmp, "transform.json"), os.path.join(tmp, "mesh_0.obj"), n_samples=ns)
        K, R, C = ref.load_cameras(os.path.join(tmp, "transform.json"), info["names"])
        out["K"], out["R"], out["C"] = K, R, C
        ref.scale_and_save_albedos(os.path.join(tmp, "albedos"), os.path.join(tmp, "scaled"), out["ratios_seed3_n400"])
        h = hashlib.sha256()
        for n in info["names"]:
            h.update(open(os.path.join(tmp, "scaled", n), "rb").read())
        import cv2
        out["scaled_view1"] = cv2.imread(os.path.join(tmp, "scaled", info["names"][1]), cv2.IMREAD_UNCHANGED)
        out["scaled_sha256"] = np.frombuffer(h.digest(), np.uint8)
        out["input_sha256"] = np.frombuffer(bytes.fromhex(info["sha256"]), np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_albedo_scaling.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    print(out["ratios_seed3_n400"])


if __name__ == "__main__":
    main()
