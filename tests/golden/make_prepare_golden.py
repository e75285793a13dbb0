#!/usr/bin/env python
"""Generates tests/golden/ref_prepare.npz by running the REFERENCE's own rnb_neus2/prepare.py (+ scaling.py), imported from /root/reference,
on the inputs of tests/albedo_scene.write_prepare_inputs for every scaling mode: the transform.json text, the sha256 of every PNG written
and the returned scaling.  Run in the build container; the fixture travels."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from albedo_scene import write_prepare_inputs  # noqa: E402

MODES = ("auto", "silhouettes", "pcd", "cameras", "none")


class Log:
    def __init__(self): self.lines = []
    def info(self, m): self.lines.append("I " + m)
    def warning(self, m): self.lines.append("W " + m)


def run(prepare_fn, tmp, mode):
    data = write_prepare_inputs(tmp)
    out = os.path.join(tmp, "out_" + mode); log = Log()
    ret = prepare_fn(data, out, log, scaling_mode=mode, sphere_scale=1.0, margin_px=20)
    files = {}
    for sub in ("normals", "albedos"):
        for n in sorted(os.listdir(os.path.join(out, sub))):
            files[sub + "/" + n] = hashlib.sha256(open(os.path.join(out, sub, n), "rb").read()).hexdigest()
    text = open(os.path.join(out, "transform.json")).read()
    log_lines = [l.replace(tmp, "<tmp>") for l in log.lines]
    return {"transform": text, "files": files, "log": log_lines, "center": np.asarray(ret["scene_center"], np.float64).tolist(), "factor": float(ret["scale_factor"]),
            "matrix": np.asarray(ret["scale_matrix"], np.float64).tolist(), "n2w": np.asarray(ret["n2w"], np.float64).tolist(), "n_frames": int(ret["n_frames"])}


def main():
    sys.path.insert(0, "/root/reference")
    from rnb_neus2 import prepare as ref
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mode in MODES:
            gold[mode] = run(ref.prepare_testbed_data, tmp, mode)
            print(mode, gold[mode]["factor"], gold[mode]["center"], gold[mode]["n_frames"])
    np.savez_compressed(os.path.join(HERE, "ref_prepare.npz"), gold=np.frombuffer(json.dumps(gold).encode(), np.uint8))


if __name__ == "__main__":
    main()
