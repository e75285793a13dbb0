#!/usr/bin/env python
"""Stage-1 -> stage-2 hand-off, step by step (GPU box): the reference build (oracle/_ref/bin/ref_harness, Testbed::load_snapshot + Testbed::train)
and the library (Testbed.load_snapshot + rnb_train) resume from the SAME reference-written snapshot with the same flags; per-step loss, rays and
sample counts of the first steps are recorded side by side (with and without --opti-lights).  EVIDENCE TOOLING.
usage: python tools/resume_probe.py OUT_DIR [--train-steps 1500] [--resume-steps 60]"""
import argparse
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader  # noqa: E402
import ref_scene  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle/_ref/bin/ref_harness")
BASE = os.path.join(ROOT, "oracle/_ref/configs/nerf/base.json")


def harness(args, log):
    with open(log, "w") as f:
        rc = subprocess.run([HARNESS] + args, stdout=f, stderr=subprocess.STDOUT, timeout=900).returncode
    rows = []
    for line in open(log, errors="replace"):
        m = re.search(r"ref step (\d+) rays (\d+) samples (\d+) compacted (\d+) loss ([0-9.eE+-]+|nan|inf)", line)
        if m:
            rows.append(dict(k=int(m.group(1)), rays=int(m.group(2)), samples=int(m.group(3)), compacted=int(m.group(4)), loss=float(m.group(5))))
    return rc, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out"); ap.add_argument("--train-steps", type=int, default=1500); ap.add_argument("--resume-steps", type=int, default=60)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    work = os.path.join(a.out, "work"); os.makedirs(work, exist_ok=True)
    pkg = rnb_loader.load_package(); scene = rnb_loader.load_scene()
    views = scene.make_scene(24, 400, 300, with_albedo=False)
    sd = os.path.join(work, "scene"); ref_scene.write_scene(sd, views, workers=8)
    snap_file = os.path.join(work, "stage1.msgpack")
    rec = {}
    rc, rows = harness([sd + "/", BASE, work, str(a.train_steps), "--no-albedo", "--time-only", "--save-snapshot", snap_file, "--print-every", "100"], os.path.join(a.out, "ref_stage1.log"))
    rec["ref_stage1"] = dict(rc=rc, tail=rows[-3:])
    for tag, extra in (("opti", ["--opti-lights"]), ("plain", [])):
        rc, rows = harness([sd + "/", BASE, work, str(a.resume_steps), "--no-albedo", "--time-only", "--load-snapshot", snap_file, "--print-every", "1"] + extra, os.path.join(a.out, "ref_resume_%s.log" % tag))
        rec["ref_resume_" + tag] = dict(rc=rc, steps=rows)
        t = pkg.Testbed(pkg.default_config(pin_rays_per_batch=0), pkg.default_flags(no_albedo=1, light_opti=1 if extra else 0, light_mode=-2))
        t.load_training_data_dir(sd)
        t.load_snapshot(snap_file)
        ours = []
        for k in range(a.resume_steps):
            st = t.train()
            ours.append(dict(k=k, rays=int(st.n_rays), samples=int(st.n_samples), compacted=int(st.n_samples_compacted), loss=float(st.loss), ek=float(st.ek_loss), mask=float(st.mask_loss), prep=int(st.density_grid_updated)))
        rec["rnb_resume_" + tag] = dict(steps=ours)
        t.close()
        r = rec["ref_resume_" + tag]["steps"]
        print(tag, "ref  loss", [round(x["loss"], 5) for x in r[:12]], "... mean(last 20)", float(np.mean([x["loss"] for x in r[-20:]])) if r else None)
        print(tag, "ours loss", [round(x["loss"], 5) for x in ours[:12]], "... mean(last 20)", float(np.mean([x["loss"] for x in ours[-20:]])))
        print(tag, "ref  rays/samples/compacted", [(x["rays"], x["samples"], x["compacted"]) for x in r[:4]])
        print(tag, "ours rays/samples/compacted", [(x["rays"], x["samples"], x["compacted"]) for x in ours[:4]], flush=True)
    json.dump(rec, open(os.path.join(a.out, "resume_probe.json"), "w"), indent=1)
    subprocess.run(["rm", "-rf", work])


if __name__ == "__main__":
    main()
