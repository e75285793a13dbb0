#!/usr/bin/env python
"""Pin the oracle (and the CUDA path) against the UNMODIFIED reference, run on the GPU box.  TEST INFRASTRUCTURE.

    python tests/ref_pin.py --config small|full --steps 6 --out gpurun_out/refpin_small [--albedo]

1. writes a synthetic scene in the reference's on-disk format (tests/ref_scene.py);
2. runs oracle/_ref/bin/ref_harness (the reference's own Testbed::train built by oracle/Makefile.ref) with the light draw
   pinned to ray_idx % 3 and rays/step pinned to 256 (no arrival-order truncation), dumping the state before and the
   outputs after every step;
3. replays every step from the dumped in-state on (a) the CPU oracle and (b) this repo's CUDA path through the C ABI and
   compares sample counts, per-ray losses, the gradient buffer, the parameters after Adam, the EMA weights and the
   occupancy grid of the next step;
4. writes summary.json (+ for the small config: golden_*.npz fixtures for tests/golden/).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader                                                   # noqa: E402
from oracle_binding import Oracle, default_flags                    # noqa: E402
from common import SMALL, FULL, product_config, copy_flags, rel_err   # noqa: E402
import ref_scene                                                    # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")


def h2f(a):
    return np.asarray(a).view(np.float16).astype(np.float32)


def groups(o):
    return {"sdf_mlp": slice(o.off_sdf, o.off_rgb), "rgb_mlp": slice(o.off_rgb, o.off_grid), "grid": slice(o.off_grid, o.off_var), "variance": slice(o.off_var, o.off_var + 1)}


def cmp_groups(o, a, b):
    return {k: rel_err(a[s], b[s]) for k, s in groups(o).items()}


def sorted_cmp(a, b):
    a = np.sort(np.asarray(a, np.float64)); b = np.sort(np.asarray(b, np.float64))
    if a.size != b.size:
        return {"n_ours": int(a.size), "n_ref": int(b.size), "rel": None}
    return {"n": int(a.size), "rel": float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)), "sum_ours": float(a.sum()), "sum_ref": float(b.sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--steps", type=int, default=34)
    ap.add_argument("--dump-steps", default="0,1,2,3,32,33")
    ap.add_argument("--rays", type=int, default=256)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--albedo", action="store_true")
    ap.add_argument("--out", default="gpurun_out/refpin")
    ap.add_argument("--work", default="/tmp/refpin")
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 4)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    work = args.work + "_" + args.config + ("_alb" if args.albedo else "")
    scene_dir = os.path.join(work, "scene"); dump = os.path.join(work, "dump")
    os.makedirs(dump, exist_ok=True)
    scene = rnb_loader.load_scene()
    views0 = scene.make_scene(args.views, args.res, args.res, with_albedo=True)
    ref_scene.write_scene(scene_dir, views0)
    cfgd = SMALL if args.config == "small" else FULL
    if args.config == "small":
        net_cfg = ref_scene.small_network_config(os.path.join(work, "small.json"))
    else:
        net_cfg = os.path.join(ROOT, "oracle", "_ref", "configs", "nerf", "base.json")
    dsteps = sorted(int(x) for x in args.dump_steps.split(",") if int(x) < args.steps)
    cmd = [HARNESS, scene_dir + "/", net_cfg, dump, str(args.steps), "--pin-rays", str(args.rays), "--dump-steps", ",".join(map(str, dsteps))]
    if not args.albedo:
        cmd.append("--no-albedo")
    t0 = time.time()
    log = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(os.path.join(args.out, "ref_harness.log"), "w").write(log.stdout)
    if log.returncode != 0:
        print(log.stdout[-3000:]); raise SystemExit("ref_harness failed rc=%d" % log.returncode)
    # second, identical run of the reference: its own run-to-run spread (atomicAdd order, fp16 atomics) is the noise floor
    dump2 = dump + "2"; os.makedirs(dump2, exist_ok=True)
    cmd2 = list(cmd); cmd2[3] = dump2
    subprocess.run(cmd2, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    meta = ref_scene.read_meta(os.path.join(dump, "meta.txt"))
    views = ref_scene.load_dataset_dump(os.path.join(dump, "dataset.bin"))
    summary = {"config": args.config, "albedo": bool(args.albedo), "rays_per_step": args.rays, "ref_meta": meta, "ref_seconds": round(time.time() - t0, 1), "steps": []}
    # the loader must hand back exactly the pixels we wrote
    summary["dataset_roundtrip"] = {"pixels_equal": bool(all(np.array_equal(a["normal"], b["normal"]) and np.array_equal(a["albedo"], b["albedo"]) for a, b in zip(views0, views))),
                                    "xform_max_abs": float(max(np.abs(np.asarray(a["xform"]) - b["xform"]).max() for a, b in zip(views0, views))),
                                    "focal_max_abs": float(max(abs(a["fx"] - b["fx"]) for a, b in zip(views0, views)))}

    def rd(name, dt):
        return np.fromfile(os.path.join(dump, name), dt)

    flags = default_flags(no_albedo=0 if args.albedo else 1, light_mode=-2, mask_loss_weight=float(meta["mask_loss_weight"]), ek_loss_weight=float(meta["ek_loss_weight"]))
    o = Oracle(threads=args.threads, **cfgd)
    assert o.n_params == int(meta["n_params"]), (o.n_params, meta["n_params"])
    o.set_flags(flags); o.set_views(views)
    t = None
    if not args.no_cuda:
        pkg = rnb_loader.load_package()
        t = pkg.Testbed(product_config(pkg, cfgd, rays_per_batch=args.rays, pin_rays_per_batch=1))
        t.set_flags(copy_flags(pkg, flags)); t.load_training_data(views)

    # --- initialisation parity: the reference's own initial parameters against rnb_init_params / the oracle's ---
    p0 = rd("step0_in_params_fp32.bin", np.float32)
    sdf_init_file = os.path.join(ROOT, "oracle", "_ref", "utils", "mlp_weights_hidden_layer_num_1_hidden_size_32.txt" if o.sdf_in == 32 else "mlp_weights.txt")
    sdf_init = np.array(open(sdf_init_file).read().split(), np.float32)
    o.init_params(1337, sdf_init)
    summary["init_params"] = {"oracle_vs_ref_max_abs": float(np.abs(o.get_params()[0] - p0).max()), "oracle_vs_ref_groups": cmp_groups(o, o.get_params()[0], p0)}
    if t is not None:
        t.init_params(sdf_init)
        summary["init_params"]["cuda_vs_ref_max_abs"] = float(np.abs(t.get_params() - p0).max())

    golden = {}
    for k in dsteps:
        tag = "step%d" % k
        pin = rd(tag + "_in_params_fp32.bin", np.float32)
        st = rd(tag + "_in_state.bin", np.uint64)
        dg = np.zeros(128 ** 3, np.float32); d_in = rd(tag + "_in_density_grid.bin", np.float32); dg[:min(d_in.size, dg.size)] = d_in[:dg.size]   # empty before the first refresh
        bf = np.zeros(128 ** 3, np.uint8); b_in = rd(tag + "_in_bitfield.bin", np.uint8); bf[:min(b_in.size, bf.size)] = b_in[:bf.size]
        cnt = rd(tag + "_out_counters.bin", np.uint64)
        R, n_a, n_b, rays_next = int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3])
        ref_loss = rd(tag + "_out_loss.bin", np.float32); ref_ek = rd(tag + "_out_ek_loss.bin", np.float32); ref_mask = rd(tag + "_out_mask_loss.bin", np.float32)
        ref_g = h2f(rd(tag + "_out_grads_fp16.bin", np.uint16))
        ref_p = rd(tag + "_out_params_fp32.bin", np.float32)
        ref_ema = h2f(rd(tag + "_out_params_ema_fp16.bin", np.uint16))
        row = {"step": k, "training_step": int(st[4]), "rays": R, "ref_samples": n_a, "ref_compacted": n_b, "ref_rays_next": rays_next}
        try:
            g2 = h2f(np.fromfile(os.path.join(dump2, tag + "_out_grads_fp16.bin"), np.uint16)); p2 = np.fromfile(os.path.join(dump2, tag + "_in_params_fp32.bin"), np.float32)
            row["ref_run_to_run"] = {"grads": cmp_groups(o, g2, ref_g), "in_params": cmp_groups(o, p2, pin),
                                     "loss_sorted": sorted_cmp(np.fromfile(os.path.join(dump2, tag + "_out_loss.bin"), np.float32), ref_loss)}
        except Exception as e:
            row["ref_run_to_run"] = {"error": str(e)}
        if k == 0:
            # Adam / EMA in isolation: the reference's own fp16 gradients through the oracle's optimizer from the same start (zero moments)
            oa = Oracle(threads=1, **cfgd); oa.set_params(pin); oa.set_grads(ref_g); oa.optimizer_step()
            pa, _, ea = oa.get_params()
            row["adam_on_ref_grads"] = {"params_max_abs": float(np.abs(pa - ref_p).max()), "params_update_rel": cmp_groups(o, pa - pin, ref_p - pin), "ema_max_abs": float(np.abs(ea - ref_ema).max())}
            if t is not None:
                pkg2 = rnb_loader.load_package(); ta = pkg2.Testbed(product_config(pkg2, cfgd, rays_per_batch=args.rays, pin_rays_per_batch=1))
                ta.set_params(pin); ta.stage_optimizer(ref_g); pc_ = ta.get_params()
                row["adam_on_ref_grads"]["cuda_params_max_abs"] = float(np.abs(pc_ - ref_p).max())
                row["adam_on_ref_grads"]["cuda_ema_max_abs"] = float(np.abs(h2f(ta.export_params_fp16(use_ema=True)) - ref_ema).max())
                ta.close()
        impls = [("oracle", o)] + ([("cuda", t)] if t is not None else [])
        for name, impl in impls:
            impl.set_params(pin)
            if name == "oracle":
                impl.set_density_grid(dg, int(st[8])); impl.set_bitfield(bf)
                impl.set_train_state(training_step=int(st[4]), rays_per_batch=R, n_rays_total=int(st[6]), measured_before=0, pin_rays=1)
                impl.set_rng(int(st[0]), int(st[1]), int(st[2]), int(st[3]))
                t1 = time.time(); s = impl.train_step(); dt = time.time() - t1
                g = impl.get_grads(); ri, lo, ek, ml = impl.last_losses()
                p_after, _, ema_after = impl.get_params()
                res = {"samples": int(s.n_samples), "compacted": int(s.n_compacted), "seconds": round(dt, 2)}
                dens_after = impl.get_density_grid(); bits_after = impl.get_bitfield()
            else:
                impl.import_density_grid(dg, int(st[8])); impl.set_bitfield(bf)
                impl.set_train_state(int(st[4]), R, int(st[6]), 0)
                impl.set_rng([int(st[0]), int(st[1]), int(st[2]), int(st[3])])
                ts = int(st[4]); skip = min(max(ts // 16, 1), 16)
                if ts % skip == 0:
                    impl.training_prep_nerf()
                impl.train_step_begin()
                g = impl.get_grads(); ri, l3 = impl.ray_losses(); lo, ek, ml = l3[:, 0], l3[:, 1], l3[:, 2]
                s = impl.train_step_end()
                p_after = impl.get_params(); ema_after = h2f(impl.export_params_fp16(use_ema=True))
                res = {"samples": int(s.n_samples), "compacted": int(s.n_samples_compacted)}
                dens_after, _ = impl.export_density_grid(); bits_after = impl.get_bitfield()
            kk = len(lo)
            res["samples_equal"] = res["samples"] == n_a
            res["compacted_equal"] = res["compacted"] == n_b
            res["loss_sorted"] = sorted_cmp(lo, ref_loss[:kk]) if np.count_nonzero(ref_loss[kk:]) == 0 else sorted_cmp(lo, ref_loss[ref_loss != 0])
            res["ek_sorted"] = sorted_cmp(ek, ref_ek[:kk])
            res["mask_sorted"] = sorted_cmp(ml, ref_mask[:kk])
            res["grads_vs_ref"] = cmp_groups(o, g, ref_g)
            res["grad_norms_ref"] = {kx: float(np.linalg.norm(ref_g[sx])) for kx, sx in groups(o).items()}
            res["params_after_vs_ref"] = cmp_groups(o, p_after - pin, ref_p - pin)       # relative error of the UPDATE
            res["params_after_max_abs"] = float(np.abs(p_after - ref_p).max())
            res["ema_after_max_abs"] = float(np.abs(np.asarray(ema_after) - ref_ema).max())
            if (k + 1) in dsteps:
                dg_next = np.zeros(128 ** 3, np.float32); d_n = rd("step%d_in_density_grid.bin" % (k + 1), np.float32); dg_next[:min(d_n.size, dg_next.size)] = d_n[:dg_next.size]
                b_next = rd("step%d_in_bitfield.bin" % (k + 1), np.uint8)
                res["density_grid_rel"] = rel_err(np.maximum(dens_after, 0), np.maximum(dg_next, 0))
                res["density_sign_mismatch"] = int(np.count_nonzero((dens_after < 0) != (dg_next < 0)))
                nb = min(b_next.size, 128 ** 3 // 8)
                res["bitfield_mip0_bits_differ"] = int(np.unpackbits(np.bitwise_xor(np.asarray(bits_after)[:nb], b_next[:nb])).sum())
                res["bitfield_mip0_bits_set_ref"] = int(np.unpackbits(b_next[:nb]).sum())
            row[name] = res
            if name == "oracle" and args.config == "small" and k in (0, dsteps[-1]):
                nm = o.off_grid
                gd = dict(state=st, ref_counters=cnt, ref_loss_sorted=np.sort(ref_loss[:kk]), ref_ek_sorted=np.sort(ref_ek[:kk]), ref_mask_sorted=np.sort(ref_mask[:kk]),
                          ref_grads_fp16=rd(tag + "_out_grads_fp16.bin", np.uint16), ref_params_out_mlp=ref_p[:nm].copy(), ref_params_out_var=ref_p[o.off_var:].copy(),
                          ref_ema_out_mlp_fp16=rd(tag + "_out_params_ema_fp16.bin", np.uint16)[:nm].copy())
                if k == 0:
                    rs = np.random.RandomState(7); idx = np.sort(rs.choice(np.arange(o.off_grid, o.off_var), 4096, replace=False))
                    gd.update(params_in_mlp=pin[:nm].copy(), params_in_grid_idx=idx.astype(np.uint32), params_in_grid_val=pin[idx].copy(),
                              params_in_grid_sum=np.array([pin[o.off_grid:o.off_var].astype(np.float64).sum(), (pin[o.off_grid:o.off_var].astype(np.float64) ** 2).sum()]))
                else:
                    gd.update(params_in_fp16=pin.astype(np.float16), bitfield=bf.copy())
                golden[tag] = gd
        print(json.dumps(row)); sys.stdout.flush()
        summary["steps"].append(row)

    # network probe: NerfNetwork::inference_mixed_precision on the final training weights
    pc = rd("probe_coords.bin", np.float32).reshape(-1, 7)
    pout = h2f(rd("probe_out_fp16.bin", np.uint16)).reshape(-1, 16)
    pf = rd("final_params_fp32.bin", np.float32); stf = rd("final_state.bin", np.uint64)
    o.set_params(pf)
    vl = o.valid_level(int(stf[4]))
    oo, _ = o.network_forward(pc, vl)
    probe = {"valid_level": vl, "oracle_vs_ref": {"albedo_raw": rel_err(oo[:, 0:3], pout[:, 0:3]), "sdf": rel_err(oo[:, 3], pout[:, 3]), "normal": rel_err(oo[:, 4:7], pout[:, 4:7]), "variance": rel_err(oo[:, 7], pout[:, 7])},
             "sdf_max_abs": float(np.abs(oo[:, 3] - pout[:, 3]).max())}
    if t is not None:
        t.set_params(pf); t.set_train_state(int(stf[4]), args.rays, int(stf[6]), int(stf[7]))
        co, _ = t.stage_forward(pc)
        probe["cuda_vs_ref"] = {"albedo_raw": rel_err(co[:, 0:3], pout[:, 0:3]), "sdf": rel_err(co[:, 3], pout[:, 3]), "normal": rel_err(co[:, 4:7], pout[:, 4:7]), "variance": rel_err(co[:, 7], pout[:, 7])}
    summary["probe"] = probe
    print(json.dumps({"probe": probe}))
    if args.config == "full":
        # sparse probe fixture for the default network (tests/golden/ref_full_probe.npz): the reference's outputs at 1024 probe points together with the
        # MLP weights and ONLY the hash entries those points read (found as the non-zero entries of a backward pass with a constant output gradient)
        npb = 1024
        dprobe = np.zeros((npb, 16), np.float32); dprobe[:, :11] = 0.01
        gtouch = o.network_backward(pc[:npb], dprobe, npb, 1 << 18, vl)
        sel = np.nonzero(gtouch[o.off_grid:o.off_var])[0]
        sel = np.unique(np.concatenate([sel & ~np.int64(1), sel | 1])).astype(np.uint32)          # whole entries (2 features)
        np.savez_compressed(os.path.join(args.out, "golden_full_probe.npz"), coords=pc[:npb].astype(np.float32), ref_out_fp16=rd("probe_out_fp16.bin", np.uint16).reshape(-1, 16)[:npb],
                            mlp_fp16=pf[:o.off_grid].astype(np.float16), var=pf[o.off_var:].astype(np.float32), grid_idx=sel, grid_val_fp16=pf[o.off_grid + sel.astype(np.int64)].astype(np.float16),
                            state=stf, valid_level=np.array([vl], np.uint32))
        print(json.dumps({"full_probe_fixture": {"points": npb, "grid_params_kept": int(sel.size), "valid_level": int(vl)}}))
    json.dump(summary, open(os.path.join(args.out, "summary_%s%s.json" % (args.config, "_alb" if args.albedo else "")), "w"), indent=1)
    if golden:
        flat = {}
        for tg, d in golden.items():
            for kx, v in d.items():
                flat[tg + "__" + kx] = v
        flat["probe_coords"] = pc[:512].astype(np.float32); flat["probe_out_fp16"] = rd("probe_out_fp16.bin", np.uint16).reshape(-1, 16)[:512]
        flat["probe_params"] = pf.astype(np.float16); flat["probe_state"] = stf
        np.savez_compressed(os.path.join(args.out, "golden_%s%s.npz" % (args.config, "_alb" if args.albedo else "")), **flat)


if __name__ == "__main__":
    main()
