#!/usr/bin/env python
"""bench.py — training rays/s of the NeuS2 / RNb-NeuS2 inner loop on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path through the C ABI
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU restatement of the reference on the host cores

Workload (N=1): BASELINE.json configs[1] — 96 views 1600x1200, normals only (--no-albedo), shipped default network
(L=14, F=2, T=2^19, SDF MLP 1x64, colour MLP 2x64), 4096 rays/step pinned.  DiLiGenT-MV is not available offline, so
the views are a synthetic analytic ellipsoid rendered on the host (rnb-neus2_b200/scene.py) and uploaded once, like the
reference's nerf_loader does.  One "step" = Testbed::train: occupancy-grid refresh when due + sample generation +
two network passes + loss + backward + Adam/EMA.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "training_rays_per_second"
UNIT = "rays/s"
RAYS_PER_STEP = 4096
WORKLOAD = "synthetic-ellipsoid 96 views 1600x1200 normals-only, hashgrid L=14 T=2^19 F=2, SDF MLP 1x64, colour MLP 2x64, 4096 rays/step pinned"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("bf16_tflops_sustained", 1400.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1400.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.rows = []; self.p = None; self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def build_views(n_views, w, h, with_albedo):
    import rnb_loader
    scene = rnb_loader.load_scene()
    from concurrent.futures import ProcessPoolExecutor
    focal = 1.37 * w
    poses = scene.camera_ring(n_views)
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
        res = list(ex.map(_render, [(p, w, h, focal, with_albedo) for p in poses]))
    views = [dict(normal=nm, albedo=al, fx=focal, fy=focal, cx=0.5, cy=0.5, xform=xf, w=w, h=h) for (nm, al), xf in zip(res, poses)]
    return views, time.time() - t0


def _render(a):
    import rnb_loader
    return rnb_loader.load_scene().render_view(*a)


def cpu_baseline_from_state(t, views, flags_kw, threads, steps, rays):
    """Oracle (CPU port of the reference algorithm) timed on the host cores from the SAME training state."""
    from oracle_binding import Oracle, default_flags
    from common import FULL
    o = Oracle(threads=threads, **FULL)
    o.set_params(t.get_params())
    o.set_views(views)
    o.set_flags(default_flags(**flags_kw))
    g, ema_step = t.export_density_grid()
    o.set_density_grid(g, ema_step); o.set_bitfield(t.get_bitfield())
    ts, _, nrt, mb = t.get_train_state()
    # a step index that does not trigger the occupancy refresh keeps the sample bounded; rays/s scales with rays
    o.set_train_state(training_step=ts | 1, rays_per_batch=rays, n_rays_total=nrt, measured_before=0, pin_rays=1)
    o.set_rng(*t.get_rng())
    o.train_step()                                   # warm-up (thread start-up, page faults)
    t0 = time.time(); n = 0
    for _ in range(steps):
        st = o.train_step(); n += 1
    dt = time.time() - t0
    return rays * n / dt, dict(samples_per_step=int(st.n_samples), compacted_per_step=int(st.n_compacted), seconds=round(dt, 2))


REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).returncode == 0
    except Exception:
        return False


def run_reference_cuda(args):
    """The UNMODIFIED reference (its own Testbed::train, tiny-cuda-nn kernels as sm_100 SASS, built by oracle/Makefile.ref into
    oracle/_ref/) on one GPU of this box, on the same scene / config / pinned 4096 rays per step.  The reference has no CPU
    implementation of this path (tiny-cuda-nn is CUDA only), so this — not a CPU port — is its own stock code path."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_scene
    work = "/tmp/rnb_bench_ref"
    scene_dir = os.path.join(work, "scene"); dump = os.path.join(work, "dump"); os.makedirs(dump, exist_ok=True)
    t0 = time.time()
    views, _ = build_views(args.views, args.width, args.height, False)
    ref_scene.write_scene(scene_dir, views, workers=min(16, os.cpu_count() or 4))
    scene_s = time.time() - t0
    n = args.pretrain + args.warmup + args.steps
    cmd = [REF_HARNESS, scene_dir + "/", os.path.join(ROOT, "oracle", "_ref", "configs", "nerf", "base.json"), dump, str(n), "--no-albedo", "--time-only",
           "--pin-rays", str(RAYS_PER_STEP), "--time-from", str(args.pretrain + args.warmup)]
    clocks = ClockSampler(0); clocks.start()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    clk = clocks.stop()
    if p.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/ref_harness failed: " + p.stdout[-200:].replace("\n", " ")})); return
    meta = ref_scene.read_meta(os.path.join(dump, "meta.txt"))
    val = float(meta["rays_per_second"]); ms = float(meta["timed_ms"]) / max(int(meta["timed_steps"]), 1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 1, "steps": int(meta["timed_steps"]), "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp16 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pretrain_steps": args.pretrain, "scene_write_s": round(scene_s, 1),
                       "note": "unmodified reference CUDA path (Testbed::train, tiny-cuda-nn; sm_100 SASS from oracle/Makefile.ref) on ONE B200 of this box, light draw pinned (oracle/ref_prelude.h), "
                               "rays/step pinned by overwriting the controller output; per-step CUDA events around Testbed::train incl. its occupancy refreshes; the reference has no CPU or multi-GPU path"},
            "clocks": clk,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "%s timed steps x %d rays on the GPU (the reference path is CUDA only; 1 host thread drives it)" % (meta["timed_steps"], RAYS_PER_STEP)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_reference(args):
    """--impl reference.  If the reference build (oracle/_ref, made by oracle/Makefile.ref) and a GPU are present: the reference's
    own CUDA path (run_reference_cuda).  Otherwise the CPU restatement (oracle/) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if os.path.exists(REF_HARNESS) and _has_gpu() and not args.cpu_port:
        return run_reference_cuda(args)
    from oracle_binding import Oracle, default_flags, build_oracle
    from common import FULL
    build_oracle()
    threads = os.cpu_count() or 1
    w, h, n_views = 1600, 1200, 96
    views, gen_s = build_views(n_views, w, h, False)
    o = Oracle(threads=threads, **FULL)
    o.init_params(1337, None)
    o.set_views(views)
    o.set_flags(default_flags(no_albedo=1))
    rays = 512        # bounded sample of the 4096-ray step: every stage scales linearly in rays
    o.set_train_state(training_step=1, rays_per_batch=rays, pin_rays=1)
    o.set_bitfield(_shell_bitfield())
    for _ in range(max(1, min(args.warmup, 2))):
        o.train_step()
    k = max(1, min(args.steps, 5))
    t0 = time.time()
    for _ in range(k):
        st = o.train_step()
    dt = time.time() - t0
    val = rays * k / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": dt / k * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 storage / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference algorithm (oracle/), occupancy = analytic surface shell, bounded sample"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": "%d rays/step x %d steps (of 4096 rays/step), %d samples/step" % (rays, k, st.n_samples)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _shell_bitfield():
    """Occupancy of a converged run: cells within ~2 cells of the analytic surface (8 mips, Morton order)."""
    import numpy as np
    import rnb_loader
    scene = rnb_loader.load_scene()
    g = (np.arange(128, dtype=np.float32) + 0.5) / 128
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt(((X - 0.5) / scene.AXES[0]) ** 2 + ((Y - 0.5) / scene.AXES[1]) ** 2 + ((Z - 0.5) / scene.AXES[2]) ** 2)
    occ = np.abs(r - 1.0) < 0.08

    def part(v):
        v = v.astype(np.uint32)
        v = (v * 0x00010001) & 0xFF0000FF; v = (v * 0x00000101) & 0x0F00F00F; v = (v * 0x00000011) & 0xC30C30C3; v = (v * 0x00000005) & 0x49249249
        return v
    ii = np.arange(128, dtype=np.uint32)
    I, J, K = np.meshgrid(ii, ii, ii, indexing="ij")
    m = (part(I) | (part(J) << 1) | (part(K) << 2)).ravel()
    bits = np.zeros(128 ** 3, np.uint8); bits[m] = occ.ravel()
    out = np.zeros(128 ** 3, np.uint8)
    out[:128 ** 3 // 8] = np.packbits(bits, bitorder="little")
    return out


DP_MODE_DEFAULT = "allreduce"      # N > 1: "allreduce" (replicated Adam) or "sharded" (reduce-scatter + sharded Adam + all-gather); RNB_DP overrides


def _network_path():
    # mirrors the selection in csrc/rnb_api.cu (RNB_NETWORK=simt|mma, RNB_BACKWARD=mma are cross-check paths)
    net = os.environ.get("RNB_NETWORK", "")
    if net in ("simt", "mma"):
        return {"simt": "CUDA-core kernels (cross-check path)", "mma": "mma.sync tile kernels (cross-check path)"}[net]
    if os.environ.get("RNB_BACKWARD", "") == "mma":
        return "tcgen05 forward + mma.sync backward"
    return "tcgen05 forward + tcgen05 backward"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--pretrain", type=int, default=300, help="untimed training steps before warm-up so that the occupancy grid is past its 256-step bootstrap (BASELINE.md §3)")
    ap.add_argument("--views", type=int, default=96)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="--impl reference: time the CPU restatement instead of the reference CUDA build")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # stdout carries exactly one JSON line: everything libraries print (NCCL banner, torch warnings) goes to stderr until then
    sys.stdout.flush()
    _saved_stdout = os.dup(1); os.dup2(2, 1)
    import numpy as np
    import torch
    import rnb_loader
    pkg = rnb_loader.load_package()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("RNB_NCCL_DEBUG", "NONE")      # keep stdout to the single JSON line (NCCL prints its version banner there otherwise)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    R = RAYS_PER_STEP * n_gpus                      # weak scaling: 4096 rays per GPU per step, global batch R
    views, gen_s = build_views(args.views, args.width, args.height, False)
    # weak scaling with the per-GPU work fixed: 4096 rays AND a 2^18-sample training batch per GPU (the reference's target_batch_size is a
    # global cap; leaving it at 2^18 for N GPUs would shrink every rank's network passes by N and read as super-linear scaling)
    cfg = pkg.default_config(rays_per_batch=R, pin_rays_per_batch=1, world_size=world, rank=rank, target_batch_size=(1 << 18) * n_gpus)
    flags_kw = dict(no_albedo=1)
    t = pkg.Testbed(cfg, pkg.default_flags(**flags_kw))
    t.init_params()
    t0 = time.time(); t.load_training_data(views); upload_s = time.time() - t0
    dataset_bytes = sum(v["normal"].nbytes for v in views)

    grad_t = stat_t = None
    dp_mode = os.environ.get("RNB_DP", DP_MODE_DEFAULT) if world > 1 else "single"
    if world > 1:
        gp, gn = t.grad_buffer(); sp, sn = t.stat_buffer()

        class _Arr:
            def __init__(self, p, n, ts="<f4"): self.__cuda_array_interface__ = {"shape": (n,), "typestr": ts, "data": (p, False), "version": 3}
        stat_t = torch.as_tensor(_Arr(sp, sn), device="cuda")
        if dp_mode == "sharded":
            # sharded optimizer: reduce-scatter of the fp32 gradients, Adam/EMA on this rank's 1/N of the parameters, all-gather
            # of the binary16 training parameters.  The arrays are padded to a multiple of 512 elements, so the shards are equal.
            pp, _, _, npad = t.param_buffers()
            shard = npad // world
            assert shard * world == npad and shard % 8 == 0
            grad_t = torch.as_tensor(_Arr(gp, npad), device="cuda")
            par_t = torch.as_tensor(_Arr(pp, npad, "<f2"), device="cuda")
            red_t = torch.zeros(shard, dtype=torch.float32, device="cuda")
            own_t = torch.zeros(shard, dtype=torch.float16, device="cuda")
            t.set_optimizer_shard(rank * shard, (rank + 1) * shard, red_t.data_ptr())
        else:
            grad_t = torch.as_tensor(_Arr(gp, gn), device="cuda")

    def step(want_stats):
        if world == 1:
            return t.train(want_stats=want_stats)
        ts = t.get_train_state()[0]
        skip = min(max(ts // 16, 1), 16)
        if ts % skip == 0:
            t.training_prep_nerf()
        t.train_step_begin()
        if dp_mode == "sharded":
            dist.reduce_scatter_tensor(red_t, grad_t); dist.all_reduce(stat_t)
            st = t.train_step_end()
            own_t.copy_(par_t[rank * shard:(rank + 1) * shard])
            dist.all_gather_into_tensor(par_t, own_t)         # every rank's next forward sees all updated shards
            return st
        dist.all_reduce(grad_t); dist.all_reduce(stat_t)      # the single gradient exchange of the step (NCCL over NVLink)
        return t.train_step_end()

    for _ in range(args.pretrain + args.warmup):
        step(False)
    torch.cuda.synchronize()

    def timed(k, want_stats):
        if dist: dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = t.launch_count()
        e0.record()
        last = None
        for _ in range(k):
            last = step(want_stats)
        e1.record()
        torch.cuda.synchronize()
        if dist: dist.barrier()
        ms = e0.elapsed_time(e1)
        if dist:
            tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms, t.launch_count() - l0, last

    # value, e2e and the per-stage pass are all taken on the SAME training steps: the state after warm-up is checkpointed on the device
    # and restored between the passes (the cost of a step changes with the training step: live hash levels, samples per ray)
    t.checkpoint_save()
    clocks = ClockSampler(local_rank); clocks.start()
    ms, launches, _ = timed(args.steps, False)
    clk = clocks.stop()
    # end to end through the public call with the per-step read-back of the loss scalars / counters
    t.checkpoint_restore()
    ms_e2e, _, last = timed(args.steps, True)
    # per-stage device timing for the roofline (events on the launching stream; separate pass so that `value` is undisturbed)
    t.checkpoint_restore()
    t.profile_enable(True)
    timed(args.steps, True)
    prof = t.profile_read(); t.profile_enable(False)

    value = R * args.steps / (ms * 1e-3)
    e2e = R * args.steps / (ms_e2e * 1e-3)
    hbm_peak, peak_src, tf_peak = peaks()
    # Per-stage device time (CUDA events on the launching stream) and algorithmic work: hash gather 8 corners x 4 B
    # per live level, scatter = read-modify-write of 8 corners x 8 B (fp32 pairs), rows in/out; MLP flops = 2 x MACs of the layers run.
    per_stage = {k: v[0] / max(v[1], 1) for k, v in prof.items()}
    ns, nc = last.n_samples, last.n_samples_trained
    ts_now = t.get_train_state()[0]
    L = int(min(14, np.ceil(0.2 * 14 + 0.02 * max(0, ts_now - 100)) + 1)) if ts_now > 0 else 14
    alg = {"march": 4.0 * ns, "scan_emit": 20.0 * ns, "pass_a_sdf_normal": (32.0 * L + 24.0) * ns, "compact": 24.0 * nc + 8.0 * ns, "pass_b_forward": (32.0 * L + 60.0) * nc, "loss": 64.0 * nc,
           "backward": (32.0 * L + 2 * 64.0 * L + 48.0) * nc, "adam_ema": 8.0 * t.n_params}
    flops = {"pass_a_sdf_normal": 8192.0 * ns, "pass_b_forward": 24576.0 * nc, "backward": 84000.0 * nc}
    stages = {}
    for k, msk in per_stage.items():
        row = {"ms": round(msk, 4)}
        if k in alg:
            row["alg_GBps"] = round(alg[k] / (msk * 1e-3) / 1e9, 1); row["hbm_frac"] = round(alg[k] / (msk * 1e-3) / 1e9 / hbm_peak, 4)
        if k in flops:
            row["TFLOPs"] = round(flops[k] / (msk * 1e-3) / 1e12, 2); row["tensor_frac"] = round(flops[k] / (msk * 1e-3) / 1e12 / tf_peak, 5)
        stages[k] = row
    dom = max((k for k in per_stage if k != "grid_update"), key=lambda k: per_stage[k])
    ach = alg.get(dom, 0.0) / (per_stage[dom] * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom)
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 storage / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_global": R, "pretrain_steps": args.pretrain, "samples_per_step": int(ns), "compacted_samples_per_step": int(nc), "live_hash_levels": L,
                       "cache": "working set (hash table 21 MB + gradients 42 MB + optimizer state 170 MB + 1.5 GB images) exceeds L2; no flush needed",
                       "parallelism": ("dp%d ray-sharded (4096 rays + 2^18-sample batch per GPU), " % n_gpus + ("fp32 gradient reduce-scatter + sharded Adam + fp16 parameter all-gather" if dp_mode == "sharded" else "fp32 gradient all-reduce")) if n_gpus > 1 else "single GPU", "network_path": _network_path(),
                       "dataset_upload_s": round(upload_s, 3), "dataset_bytes": dataset_bytes, "scene_render_s": round(gen_s, 1)},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64,
                    "note": "rnb_train through the C ABI with per-step stats read-back; dataset resident after one upload (%.2f s for %d MB from host memory), as in the reference" % (upload_s, dataset_bytes >> 20)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "achieved = algorithmic bytes of the stage / its CUDA-event time; the hash table (21 MB) and gradient buffer (42 MB) are L2 resident, so DRAM traffic (ncu) is far below the algorithmic bytes",
                         "stages": stages}}
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, info = cpu_baseline_from_state(t, views, flags_kw, threads, 3, 512)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": "3 steps x 512 rays from the same training state (%s)" % json.dumps(info)}
    sys.stdout.flush(); os.dup2(_saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line)); sys.stdout.flush()
    if dist:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
