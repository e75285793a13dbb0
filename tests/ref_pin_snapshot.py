#!/usr/bin/env python
"""Pin the snapshot hand-off against the UNMODIFIED reference on the GPU box.  TEST INFRASTRUCTURE.

    python tests/ref_pin_snapshot.py --out gpurun_out/refpin_snapshot

A. reference -> here: the reference trains the small network for a few steps and calls Testbed::save_snapshot as
   src/main.cu:468 does.  The file is decoded with rnb-neus2_b200/snapshot.py; checked: re-encoding reproduces the file byte for
   byte (same MessagePack choices as nlohmann::json), params_binary == the trainer's inference parameters, density_grid_binary ==
   binary16 of the density grid, the movement blobs == this repo's defaults for a static scene; Testbed.load_snapshot (CUDA path)
   takes it over (parameters, grid, controller state).
B. here -> reference: the CUDA path trains a few steps, writes a snapshot with Testbed.save_snapshot into the same network
   config; the reference loads it with Testbed::load_snapshot (src/main.cu:312) and dumps what it now holds: parameters, density
   grid, bitfield, controller state, and a network probe.  Compared against the writer's state and the CUDA forward.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader                                                   # noqa: E402
from common import SMALL, product_config, rel_err                   # noqa: E402
import ref_scene                                                    # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")


def h2f(a):
    return np.asarray(a, np.uint16).view(np.float16).astype(np.float32)


def structure(o):
    if isinstance(o, dict):
        return {k: structure(v) for k, v in o.items()}
    if isinstance(o, list):
        return [structure(v) for v in o[:8]] + (["... %d more" % (len(o) - 8)] if len(o) > 8 else [])
    if isinstance(o, (bytes, bytearray)):
        return {"bin": len(o), "sha256": hashlib.sha256(o).hexdigest()[:16], "head": bytes(o[:24]).hex()}
    return o


def run(cmd, log_path):
    log = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(log_path, "w").write(log.stdout[-200000:])
    if log.returncode != 0:
        print(log.stdout[-3000:]); raise SystemExit("ref_harness failed rc=%d" % log.returncode)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--rays", type=int, default=256)
    ap.add_argument("--out", default="gpurun_out/refpin_snapshot")
    ap.add_argument("--work", default="/tmp/refpin_snapshot")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    work = args.work; scene_dir = os.path.join(work, "scene"); dumpA = os.path.join(work, "dumpA"); dumpB = os.path.join(work, "dumpB")
    for d in (dumpA, dumpB):
        os.makedirs(d, exist_ok=True)
    pkg = rnb_loader.load_package()
    from rnb_neus2_b200 import snapshot as snap
    scene = rnb_loader.load_scene()
    views0 = scene.make_scene(8, 128, 128, with_albedo=True)
    ref_scene.write_scene(scene_dir, views0)
    net_cfg = ref_scene.small_network_config(os.path.join(work, "small.json"))
    summary = {}

    # ---- A: reference writes, we read ----
    ref_file = os.path.join(work, "ref.msgpack")
    run([HARNESS, scene_dir + "/", net_cfg, dumpA, str(args.steps), "--pin-rays", str(args.rays), "--time-only", "--save-snapshot", ref_file], os.path.join(args.out, "harness_A.log"))
    raw = open(ref_file, "rb").read()
    cfg = snap.unpackb(raw)
    summary["ref_file_bytes"] = len(raw)
    summary["structure"] = structure(cfg)
    summary["reencode_byte_identical"] = snap.packb(cfg) == raw
    try:
        import msgpack
        summary["msgpack_package_agrees"] = msgpack.unpackb(raw, raw=False, strict_map_key=False) == cfg
    except ImportError:
        summary["msgpack_package_agrees"] = None
    d = snap.parse_snapshot(cfg)
    p_inf = np.fromfile(os.path.join(dumpA, "snapshot_params_inference_fp16.bin"), np.uint16)
    grid = np.fromfile(os.path.join(dumpA, "snapshot_density_grid.bin"), np.float32)
    summary["A"] = {"params_binary_equal": bool(np.array_equal(d["params_fp16"].view(np.uint16), p_inf)),
                    "density_grid_binary_equal": bool(np.array_equal(d["density_grid"].astype(np.float16).view(np.uint16), grid.astype(np.float16).view(np.uint16))),
                    "training_step": d["training_step"], "rays_per_batch": d["rays_per_batch"], "measured_batch_size": d["measured_batch_size"],
                    "measured_batch_size_before_compaction": d["measured_batch_size_before_compaction"], "aabb_scale": d["aabb_scale"]}
    mv = snap.movement_defaults()
    summary["A"]["movement_defaults_equal"] = {k: bool(cfg["snapshot"].get(k) == v) for k, v in mv.items()}
    views = views0
    t = pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=args.rays, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    t.load_training_data(views)
    t.load_snapshot(ref_file)
    ts = t.get_train_state()
    g_after, _ = t.export_density_grid()
    summary["A"]["cuda_load"] = {"params_equal": bool(np.array_equal(t.export_params_fp16(use_ema=True).view(np.uint16), p_inf)),
                                 "train_params_equal": bool(np.array_equal(t.export_params_fp16(use_ema=False).view(np.uint16), p_inf)),
                                 "master_equal": bool(np.array_equal(t.get_params(), h2f(p_inf))),
                                 "grid_equal": bool(np.array_equal(g_after, d["density_grid"])), "train_state": ts}
    st = t.train()                                           # the loaded state trains on
    summary["A"]["cuda_load"]["next_step_loss_finite"] = bool(np.isfinite(st.loss))

    # ---- B: we write, the reference reads ----
    t2 = pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=args.rays, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    sdf_init = np.array(open(os.path.join(ROOT, "oracle", "_ref", "utils", "mlp_weights_hidden_layer_num_1_hidden_size_32.txt")).read().split(), np.float32)
    t2.init_params(sdf_init); t2.load_training_data(views)
    for _ in range(args.steps):
        t2.train()
    our_file = os.path.join(work, "ours.msgpack")
    base_cfg = {k: v for k, v in cfg.items() if k != "snapshot"}
    t2.save_snapshot(our_file, base_cfg)
    ours = snap.parse_snapshot(snap.read_snapshot(our_file))
    run([HARNESS, scene_dir + "/", net_cfg, dumpB, "0", "--load-snapshot", our_file], os.path.join(args.out, "harness_B.log"))
    pf = np.fromfile(os.path.join(dumpB, "final_params_fp32.bin"), np.float32)
    gB = np.fromfile(os.path.join(dumpB, "final_density_grid.bin"), np.float32)
    bB = np.fromfile(os.path.join(dumpB, "final_bitfield.bin"), np.uint8)
    stB = np.fromfile(os.path.join(dumpB, "final_state.bin"), np.uint64)
    t3 = pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=args.rays, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    t3.load_training_data(views); t3.load_snapshot(our_file)
    b3 = t3.get_bitfield(); nb = min(bB.size, 128 ** 3 // 8)
    pc = np.fromfile(os.path.join(dumpB, "probe_coords.bin"), np.float32).reshape(-1, 7)
    pout = h2f(np.fromfile(os.path.join(dumpB, "probe_out_fp16.bin"), np.uint16)).reshape(-1, 16)
    co, _ = t3.stage_forward(pc)
    # the reference has not run Testbed::train since the load, so its encoding still has every level enabled (the progressive
    # mask is set per training step, src/testbed.cu:2787-2793); the like-for-like probe on our side is training step 0
    t3.set_train_state(0, ours["rays_per_batch"], 0, 0)
    co_all, _ = t3.stage_forward(pc)
    summary["B"] = {"file_bytes": os.path.getsize(our_file),
                    "ref_params_equal": bool(np.array_equal(pf, ours["params_fp16"].astype(np.float32))),
                    "ref_grid_equal": bool(np.array_equal(gB, ours["density_grid"])),
                    "ref_training_step": int(stB[4]), "our_training_step": ours["training_step"],
                    "ref_rays_per_batch": int(stB[5]), "our_rays_per_batch": ours["rays_per_batch"],
                    "ref_measured_before": int(stB[7]), "our_measured_before": ours["measured_batch_size_before_compaction"],
                    "ref_measured": int(stB[9]), "our_measured": ours["measured_batch_size"],
                    "bitfield_mip0_bits_differ": int(np.unpackbits(np.bitwise_xor(np.asarray(b3)[:nb], bB[:nb])).sum()), "bitfield_mip0_bits_set": int(np.unpackbits(bB[:nb]).sum()),
                    "probe_cuda_vs_ref": {"albedo_raw": rel_err(co[:, 0:3], pout[:, 0:3]), "sdf": rel_err(co[:, 3], pout[:, 3]), "normal": rel_err(co[:, 4:7], pout[:, 4:7])},
                    "probe_cuda_all_levels_vs_ref": {"albedo_raw": rel_err(co_all[:, 0:3], pout[:, 0:3]), "sdf": rel_err(co_all[:, 3], pout[:, 3]), "normal": rel_err(co_all[:, 4:7], pout[:, 4:7])}}
    print(json.dumps(summary)[:6000])
    json.dump(summary, open(os.path.join(args.out, "summary_snapshot.json"), "w"), indent=1)
    # fixture for the CPU suite: the reference's file with the two big blobs cut to their first 4 KiB (+ their hashes in the summary)
    small = snap.unpackb(raw)
    for k in ("params_binary", "density_grid_binary"):
        small["snapshot"][k] = small["snapshot"][k][:4096]
    np.savez_compressed(os.path.join(args.out, "golden_snapshot_small.npz"), truncated=np.frombuffer(snap.packb(small), np.uint8),
                        params_head=p_inf[:2048], grid_head=grid[:2048], full_sha256=np.frombuffer(hashlib.sha256(raw).digest(), np.uint8))


if __name__ == "__main__":
    main()
