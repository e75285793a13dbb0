#!/usr/bin/env python
"""The drop-in binary on TWO GPUs (GPU box, `gpurun --gpus 2`): two copies of oracle/_ref/bin/testbed_rnb with the reference's own command line,
RNB_WORLD_SIZE=2 RNB_RANK=r RNB_COMM_ID_FILE=... (shim/rnb_testbed_shim.h), each on its own copy of the scene directory, next to ONE copy on one GPU.
Recorded: return codes, the `iteration= loss=` lines, whether the two ranks wrote byte-identical snapshots, wall times.  EVIDENCE TOOLING.
usage: python tools/shim_dp_run.py OUT_DIR [--steps 300]"""
import argparse, json, os, re, shutil, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader, ref_scene  # noqa: E402

BIN = os.path.join(ROOT, "oracle/_ref/bin/testbed_rnb")


def loss_lines(path):
    return [l.strip()[:120] for l in open(path, errors="replace") if re.search(r"iteration=\d+", l)][-4:]


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("out"); ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    work = "/dev/shm/rnb_shim_dp"; shutil.rmtree(work, ignore_errors=True); os.makedirs(work)
    views = rnb_loader.load_scene().make_scene(12, 320, 240, with_albedo=False)
    ref_scene.write_scene(os.path.join(work, "s0"), views, workers=4)
    for k in (1, 2):
        shutil.copytree(os.path.join(work, "s0"), os.path.join(work, "s%d" % k))
    argv = ["--no-gui", "--maxiter", str(a.steps), "--mask-weight", "1.0", "--save-snapshot", "--no-albedo"]
    rec = dict(argv=argv, steps=a.steps)
    t0 = time.time(); procs = []
    for r in (0, 1):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(r), RNB_WORLD_SIZE="2", RNB_RANK=str(r), RNB_COMM_ID_FILE=os.path.join(work, "nccl_id"))
        log = open(os.path.join(a.out, "rank%d.log" % r), "w")
        procs.append((subprocess.Popen([BIN, "--scene", os.path.join(work, "s%d" % r) + "/"] + argv, stdout=log, stderr=subprocess.STDOUT, env=env), log))
    rcs = []
    for p, log in procs:
        try:
            rcs.append(p.wait(timeout=90))
        except subprocess.TimeoutExpired:
            p.kill(); rcs.append(-9)
        log.close()
    rec["two_ranks"] = dict(rc=rcs, wall_s=round(time.time() - t0, 2), loss_lines=[loss_lines(os.path.join(a.out, "rank%d.log" % r)) for r in (0, 1)])
    snaps = [os.path.join(work, "s%d" % r, "output", "snapshot_%d.msgpack" % a.steps) for r in (0, 1, 2)]
    if all(os.path.exists(s) for s in snaps[:2]):
        b0, b1 = open(snaps[0], "rb").read(), open(snaps[1], "rb").read()
        rec["two_ranks"]["snapshots_byte_identical"] = b0 == b1; rec["two_ranks"]["snapshot_bytes"] = len(b0)
    t0 = time.time()
    with open(os.path.join(a.out, "single.log"), "w") as log:
        try:
            rc = subprocess.run([BIN, "--scene", os.path.join(work, "s2") + "/"] + argv, stdout=log, stderr=subprocess.STDOUT, env=dict(os.environ, CUDA_VISIBLE_DEVICES="0"), timeout=90).returncode
        except subprocess.TimeoutExpired:
            rc = -9
    rec["single"] = dict(rc=rc, wall_s=round(time.time() - t0, 2), loss_lines=loss_lines(os.path.join(a.out, "single.log")))
    if os.path.exists(snaps[2]) and os.path.exists(snaps[0]):
        try:
            import importlib
            import numpy as np
            rnb_loader.load_package()
            snap = importlib.import_module("rnb_neus2_b200.snapshot")
            pa = snap.parse_snapshot(snap.read_snapshot(snaps[0])); pb = snap.parse_snapshot(snap.read_snapshot(snaps[2]))
            wa, wb = pa["params_fp16"].astype(np.float64), pb["params_fp16"].astype(np.float64)
            rec["params_rel_two_ranks_vs_single"] = float(np.linalg.norm(wa - wb) / np.linalg.norm(wb))
            rec["snapshot_fields"] = dict(two_ranks={k: pa[k] for k in ("training_step", "loss", "rays_per_batch")}, single={k: pb[k] for k in ("training_step", "loss", "rays_per_batch")})
        except Exception as e:      # the comparison is a bonus; the record above stands without it
            rec["snapshot_compare_error"] = repr(e)[:200]
    json.dump(rec, open(os.path.join(a.out, "shim_dp_record.json"), "w"), indent=1)
    print(json.dumps(rec)[:1800])
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
