#!/usr/bin/env python
"""bench.py — training rays/s of the NeuS2 / RNb-NeuS2 inner loop on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path through the C ABI
    python bench.py --impl reference --gpus N --steps K --warmup W   # the unmodified reference CUDA build (oracle/_ref) on one GPU, else the CPU restatement

Workload: the north_star target — a synthetic 96-view 1600x1200 normal + albedo scene (DiLiGenT-MV is not available offline: analytic ellipsoid
with a procedural albedo, rnb-neus2_b200/scene.py), shipped default network (L=14, F=2, T=2^19, SDF MLP 1x64, colour MLP 2x64), RGB+ reflectance
loss, 4096 rays/step per GPU pinned, timed after 700 training steps so that all 14 hash levels are live (grid.h:1430-1437) and the occupancy grid
is past its 256-step bootstrap.  `--workload normals` is BASELINE configs[1] (--no-albedo).  One "step" = Testbed::train: occupancy-grid refresh
when due + sample generation + two network passes + loss + backward + (N > 1: one binary16 gradient all-reduce inside the library) + Adam/EMA.
Prints ONE JSON line on rank 0; further operating points (normals-only, the adaptive batch-size controller the drop-in binary runs,
N > 1: BASELINE configs[3] = 16,384 rays/step strong-sharded) ride in `records`.
"""
import argparse
import glob
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
NETWORK = "hashgrid L=14 T=2^19 F=2, SDF MLP 1x64, colour MLP 2x64"
WORKLOADS = {
    "albedo": "synthetic-ellipsoid 96 views 1600x1200 normals+albedo (RGB+ reflectance loss), " + NETWORK + ", 4096 rays/step pinned",
    "normals": "synthetic-ellipsoid 96 views 1600x1200 normals-only, " + NETWORK + ", 4096 rays/step pinned",
}
WORKLOAD = WORKLOADS["albedo"]
PRETRAIN_DEFAULT = 700


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
    """nvidia-smi clocks / throttle reasons during the timed region.  The sampler is started ahead of the region (on a box with 8 GPUs nvidia-smi needs longer
    to deliver its first row than 200 steps take) and every row carries its arrival time: `mark_begin()` opens the region, `stop()` closes it and keeps
    the rows inside; when the region is shorter than one sample, the rows of the warm-up steps right before it (the same kernels, the GPU never idles
    in between) stand in and `window` says so."""

    def __init__(self, index=0, enabled=True):
        self.rows = []; self.p = None; self.index = index; self.enabled = enabled; self.t_begin = None

    def start(self):
        if not self.enabled:
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        time.sleep(0.15)
        self.p.terminate()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        rows = [r for ts, r in self.rows if t0 <= ts <= t_end + 0.15]; window = None
        if not rows:
            rows = [r for ts, r in self.rows if ts >= t0 - 1.0]; window = "warm-up steps + timed region (the region is shorter than one nvidia-smi sample)"
        if not rows:
            rows = [r for _, r in self.rows]; window = "pretraining + warm-up + timed region (no later sample arrived)"
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}
        if window and sm:
            out["window"] = window
        return out


def build_views(n_views, w, h, with_albedo):
    import numpy as np
    import rnb_loader
    scene = rnb_loader.load_scene()
    from concurrent.futures import ProcessPoolExecutor
    focal = 1.37 * w
    poses = scene.camera_ring(n_views)
    t0 = time.time()
    # RNB_BENCH_CACHE=<dir>: keep the rendered maps between runs of one session (the scene is deterministic; rendering it takes ~20 s of host time)
    cache = os.environ.get("RNB_BENCH_CACHE")
    cf = os.path.join(cache, "views_%d_%d_%d_%d.npz" % (n_views, w, h, int(with_albedo))) if cache else None
    if cf and os.path.exists(cf):
        z = np.load(cf)
        res = [(z["n%d" % i], z["a%d" % i] if with_albedo else None) for i in range(n_views)]
    else:
        with ProcessPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
            res = list(ex.map(_render, [(p, w, h, focal, with_albedo) for p in poses]))
        if cf and int(os.environ.get("RANK", "0")) == 0:
            os.makedirs(cache, exist_ok=True)
            d = {"n%d" % i: r[0] for i, r in enumerate(res)}
            if with_albedo:
                d.update({"a%d" % i: r[1] for i, r in enumerate(res)})
            np.savez(cf + ".tmp.npz", **d); os.replace(cf + ".tmp.npz", cf)
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
    albedo = args.workload == "albedo"
    views, _ = build_views(args.views, args.width, args.height, albedo)
    ref_scene.write_scene(scene_dir, views, workers=min(16, os.cpu_count() or 4))
    scene_s = time.time() - t0
    n = args.pretrain + args.warmup + args.steps
    cmd = [REF_HARNESS, scene_dir + "/", os.path.join(ROOT, "oracle", "_ref", "configs", "nerf", "base.json"), dump, str(n)] + ([] if albedo else ["--no-albedo"]) + ["--time-only",
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
            "config": {"workload": WORKLOADS[args.workload], "pretrain_steps": args.pretrain, "scene_write_s": round(scene_s, 1),
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
    albedo = args.workload == "albedo"
    views, gen_s = build_views(n_views, w, h, albedo)
    o = Oracle(threads=threads, **FULL)
    o.init_params(1337, None)
    o.set_views(views)
    o.set_flags(default_flags(no_albedo=0 if albedo else 1))
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
            "config": {"workload": WORKLOADS[args.workload], "note": "CPU restatement of the reference algorithm (oracle/), occupancy = analytic surface shell, bounded sample"},
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


DP_MODE_DEFAULT = "sharded"        # N > 1: "sharded" (binary16 reduce-scatter + Adam on 1/N of the parameters + parameter all-gather) or "allreduce" (one binary16 all-reduce, replicated Adam); RNB_DP overrides


def _network_path():
    # mirrors the selection in csrc/rnb_api.cu (RNB_NETWORK=simt|mma, RNB_BACKWARD=mma are cross-check paths)
    net = os.environ.get("RNB_NETWORK", "")
    if net in ("simt", "mma"):
        return {"simt": "CUDA-core kernels (cross-check path)", "mma": "mma.sync tile kernels (cross-check path)"}[net]
    if os.environ.get("RNB_BACKWARD", "") == "mma":
        return "tcgen05 forward + mma.sync backward"
    return "tcgen05 forward + tcgen05 backward"


def live_levels(step, n_levels=14):
    """hash levels a training step touches (progressive training, grid.h:1430-1437)"""
    import math
    if step <= 0:
        return n_levels
    return int(min(n_levels, math.ceil(0.2 * n_levels + 0.02 * max(0, step - 100)) + 1))


def algorithmic_work(L, ns, nc, n_params):
    """SURVEY §8(d) / DESIGN §7, per stage: hash gather 8 corners x 4 B per live level and sample; merged first + second order scatter as a
    read-modify-write of 8 corners x 4 B (the reference's binary16 gradient format) = 64 B per level; sample rows in / out; optimizer 8 B per
    parameter (lower bound: gradient read + EMA, +34 B per touched parameter not counted).  MLP flops = 2 x MACs of the layers a stage runs."""
    alg = {"march": 4.0 * ns, "scan_emit": 20.0 * ns, "pass_a_sdf_normal": (32.0 * L + 24.0) * ns, "compact": 24.0 * nc + 8.0 * ns, "pass_b_forward": (32.0 * L + 48.0) * nc, "loss": 64.0 * nc,
           "backward": ((32.0 + 64.0) * L + 48.0) * nc, "adam_ema": 8.0 * n_params}
    flops = {"pass_a_sdf_normal": 10368.0 * ns, "pass_b_forward": 26752.0 * nc, "backward": 84000.0 * nc}
    return alg, flops


def _traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the newest committed ncu --set full capture (profiles/rNN_traffic.json)"""
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            v = json.load(open(f)).get(kernel)
            if v is not None:
                return v, os.path.basename(f)
        except Exception:
            pass
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="albedo", choices=sorted(WORKLOADS), help="albedo: the north_star target (normals + albedo, RGB+ loss); normals: BASELINE configs[1] (--no-albedo)")
    ap.add_argument("--pretrain", type=int, default=PRETRAIN_DEFAULT, help="untimed training steps before warm-up: all 14 hash levels live (step > 560), occupancy grid past its 256-step bootstrap")
    ap.add_argument("--views", type=int, default=96)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-records", action="store_true", help="only the headline workload (skip the normals-only / adaptive / strong-scaling records)")
    ap.add_argument("--cpu-port", action="store_true", help="--impl reference: time the CPU restatement instead of the reference CUDA build")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # stdout carries exactly one JSON line: everything libraries print (NCCL banner / NCCL_DEBUG output, torch warnings) goes to stderr until then
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
        import torch.distributed as dist          # plumbing only (unique-id broadcast, barriers, max over ranks); the gradient exchange is inside the library
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    albedo = args.workload == "albedo"
    views, gen_s = build_views(args.views, args.width, args.height, albedo)
    dataset_bytes = sum(v["normal"].nbytes + (v["albedo"].nbytes if v["albedo"] is not None else 0) for v in views)
    flags_kw = dict(no_albedo=0 if albedo else 1)
    dp_mode = os.environ.get("RNB_DP", DP_MODE_DEFAULT) if world > 1 else "single"
    if world > 1:
        os.environ["RNB_DP"] = dp_mode

    def make_testbed(rays_global, target_global, pin=1, fl=None):
        cfg = pkg.default_config(rays_per_batch=rays_global, pin_rays_per_batch=pin, world_size=world, rank=rank, target_batch_size=target_global)
        tb = pkg.Testbed(cfg, pkg.default_flags(**(fl or flags_kw)))
        tb.init_params()
        t0 = time.time(); tb.load_training_data(views); up = time.time() - t0
        if world > 1:
            ids = [pkg.Testbed.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, 0)
            tb.comm_init(ids[0])
        return tb, up

    # the training stream: created with a priority above the lowest, so that the library's side stream (lowest priority: the march of the next step,
    # launched one step ahead) only fills what the step's own kernels leave free (RNB_BENCH_STREAM=default: the legacy default stream)
    if os.environ.get("RNB_BENCH_STREAM", "high") == "default":
        tstream = torch.cuda.current_stream(); sh = None
    else:
        tstream = torch.cuda.Stream(priority=-1); sh = tstream.cuda_stream

    def timed(tb, k, want_stats):
        if dist: dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = tb.launch_count()
        e0.record(tstream)
        last = None; rays = 0; acc = [0, 0]
        for _ in range(k):
            last = tb.train(stream=sh, want_stats=want_stats)
            if want_stats:
                rays += int(last.n_rays); acc[0] += int(last.n_samples); acc[1] += int(last.n_samples_trained)
        timed.mean_counts = (acc[0] / k, acc[1] / k) if want_stats else None
        e1.record(tstream)
        torch.cuda.synchronize()
        if dist: dist.barrier()
        ms = e0.elapsed_time(e1)
        if dist:
            tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms, tb.launch_count() - l0, last, rays

    # ---- headline: weak scaling with the per-GPU work fixed: 4096 rays AND a 2^18-sample training batch per GPU (the reference's target_batch_size
    # is a global cap; leaving it at 2^18 for N GPUs would shrink every rank's network passes by N and read as super-linear scaling)
    R = RAYS_PER_STEP * n_gpus
    clocks = ClockSampler(local_rank, enabled=rank == 0); clocks.start()      # rank 0's line is the one printed; started here (before the upload, the communicator and the pretraining steps) so that rows are flowing when the timed region opens
    t, upload_s = make_testbed(R, (1 << 18) * n_gpus)
    for _ in range(args.pretrain + args.warmup):
        t.train(stream=sh, want_stats=False)
    torch.cuda.synchronize()
    # value, e2e and the per-stage pass are all taken on the SAME training steps: the state after warm-up is checkpointed on the device
    # and restored between the passes (the cost of a step changes with the training step: live hash levels, samples per ray)
    t.checkpoint_save()
    clocks.mark_begin()
    ms, launches, _, _ = timed(t, args.steps, False)          # no per-step read-back: rnb_train returns without host synchronisation
    clk = clocks.stop()
    t.checkpoint_restore()
    ms_e2e, _, last, _ = timed(t, args.steps, True)            # end to end through the public call with the per-step read-back of the loss scalars / counters
    t.checkpoint_restore()
    t.profile_enable(True)                                     # per-stage device timing for the roofline (events on the launching stream; separate pass so that `value` is undisturbed)
    timed(t, args.steps, True)
    mean_ns, mean_nc = timed.mean_counts          # samples before / after compaction, averaged over the SAME steps the stage times are averaged over
    prof = t.profile_read(); t.profile_enable(False)

    value = R * args.steps / (ms * 1e-3)
    e2e = R * args.steps / (ms_e2e * 1e-3)
    hbm_peak, peak_src, tf_peak = peaks()
    per_stage = {k: v[0] / max(v[1], 1) for k, v in prof.items()}
    ns, nc = mean_ns, mean_nc
    ts_now = t.get_train_state()[0]
    L = live_levels(ts_now)
    alg, flops = algorithmic_work(L, ns, nc, t.n_params)
    stages = {}
    for k, msk in per_stage.items():
        row = {"ms": round(msk, 4)}
        if k in alg:
            row["alg_GBps"] = round(alg[k] / (msk * 1e-3) / 1e9, 1); row["hbm_frac"] = round(alg[k] / (msk * 1e-3) / 1e9 / hbm_peak, 4)
        if k in flops:
            row["TFLOPs"] = round(flops[k] / (msk * 1e-3) / 1e12, 2); row["tensor_frac"] = round(flops[k] / (msk * 1e-3) / 1e12 / tf_peak, 5)
        stages[k] = row
    comm_stages = ("grid_update", "grad_pack", "grad_exchange", "param_allgather", "prefix_exchange")
    dom = max((k for k in per_stage if k not in comm_stages), key=lambda k: per_stage[k])
    ach = alg.get(dom, 0.0) / (per_stage[dom] * 1e-3) / 1e9
    traffic, traffic_src = _traffic(dom)
    step_bytes = sum(alg[k] for k in alg if k in per_stage)
    step_flops = sum(flops[k] for k in flops if k in per_stage)
    if n_gpus > 1:
        par = "dp%d ray-sharded (4096 rays + 2^18-sample batch per GPU); gradient exchange inside the library (rnb_comm_init): " % n_gpus + (
            "binary16 reduce-scatter + sharded Adam + binary16 parameter all-gather" if dp_mode == "sharded" else "ONE binary16 all-reduce of the 21 MB gradient buffer (+ 8 floats of statistics in the same NCCL group)")
    else:
        par = "single GPU"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 storage / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "rays_per_step_global": R, "pretrain_steps": args.pretrain, "training_step_at_end": int(ts_now), "samples_per_step": int(ns), "compacted_samples_per_step": int(nc), "counts_note": "mean over the timed window",
                       "live_hash_levels": L,
                       "cache": "working set (hash table 21 MB + gradients 42 MB + optimizer state 170 MB + %.1f GB images) exceeds L2; no flush needed" % (dataset_bytes / 1e9),
                       "parallelism": par, "network_path": _network_path(), "stream": "caller's stream with priority -1 (library side stream: lowest priority)" if sh is not None else "legacy default stream",
                       "dataset_upload_s": round(upload_s, 3), "dataset_bytes": dataset_bytes, "scene_render_s": round(gen_s, 1)},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64,
                    "note": "rnb_train through the C ABI with per-step stats read-back (host waits for every step); `value` = the same steps without read-back (no host synchronisation); "
                            "dataset resident after one upload (%.2f s for %d MB from host memory), as in the reference" % (upload_s, dataset_bytes >> 20)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "note": "achieved = algorithmic bytes of the stage (SURVEY 8(d): 32 B gather + 64 B scatter per sample-level, rows in/out) / its CUDA-event time; the hash table (21 MB) and "
                                 "the gradient buffer (42 MB) are L2 resident, so DRAM traffic (ncu) is far below the algorithmic bytes",
                         "step": {"alg_GBps": round(step_bytes / (ms / args.steps * 1e-3) / 1e9, 1), "hbm_frac": round(step_bytes / (ms / args.steps * 1e-3) / 1e9 / hbm_peak, 4),
                                  "TFLOPs": round(step_flops / (ms / args.steps * 1e-3) / 1e12, 2), "tensor_frac": round(step_flops / (ms / args.steps * 1e-3) / 1e12 / tf_peak, 5)},
                         "stages": stages}}
    if world > 1:
        line["config"]["nccl"] = t.comm_info()

    # ---- CPU baseline beside it: the oracle (CPU port of the reference algorithm) on all host cores from the same trained state, bounded sample
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        t.checkpoint_restore()
        threads = os.cpu_count() or 1
        cs, cr = 8, RAYS_PER_STEP      # ~10 s on 16 cores: whole steps of the workload, not a sub-sampled batch
        v, info = cpu_baseline_from_state(t, views, flags_kw, threads, cs, cr)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": "%d steps x %d rays (whole steps of the workload) from the same trained state (%s)" % (cs, cr, json.dumps(info))}

    # ---- further operating points (shorter windows) ------------------------------------------------------------------------------
    if not args.no_records:
        records = {}
        k2 = max(20, args.steps // 3)
        # (1) the other loss configuration on the same trained state and the same steps (the flag only changes the loss kernel and the colour-MLP gradient)
        if albedo:
            t.checkpoint_restore()
            t.set_flags(pkg.default_flags(no_albedo=1))
            m2, _, _, _ = timed(t, k2, False)
            t.set_flags(pkg.default_flags(**flags_kw))
            records["normals"] = {"workload": WORKLOADS["normals"], "value": R * k2 / (m2 * 1e-3), "unit": UNIT, "ms_per_step": m2 / k2, "steps": k2, "note": "same trained state and steps, --no-albedo loss"}
        # (1b) BASELINE configs[4]: --supernormal (identity light basis, normals + Eikonal) on the same state, and the mesh-resolution-1024 extraction
        # (SDF sweep over the 1024^3 lattice on the tcgen05 probe kernel + marching cubes + normals + colours; device times from CUDA events)
        t.checkpoint_restore()
        t.set_flags(pkg.default_flags(no_albedo=1, apply_supernormal=1))
        m5, _, _, _ = timed(t, k2, False)
        t.set_flags(pkg.default_flags(**flags_kw))
        records["supernormal"] = {"workload": WORKLOADS["normals"].replace("normals-only", "normals-only --supernormal"), "value": R * k2 / (m5 * 1e-3), "unit": UNIT, "ms_per_step": m5 / k2, "steps": k2}
        if n_gpus == 1:
            try:
                t.checkpoint_restore()
                torch.cuda.synchronize(); t0 = time.time()
                mi = t.marching_cubes(1024)
                torch.cuda.synchronize()
                records["mesh_1024"] = {"wall_s": round(time.time() - t0, 4), "n_vertices": int(mi["n_verts"]), "n_triangles": int(mi["n_indices"]) // 3,
                                        "stage_ms": mi["stage_ms"],
                                        "lattice_points": 1024 ** 3, "note": "Testbed::marching_cubes at --resolution 1024 on the trained state (EMA weights)"}
            except Exception as ex:      # noqa: BLE001
                records["mesh_1024"] = {"error": str(ex)[:200]}
        # (2) the adaptive batch-size controller (pin_rays_per_batch = 0): what shim/rnb_testbed_shim.h and therefore ./build/testbed run; the controller holds
        # the compacted sample count at 2^18 per GPU, rays/step follow the scene; the host waits for the counters every step like the reference
        ta, _ = make_testbed(R, (1 << 18) * n_gpus, pin=0)
        for _ in range(args.pretrain + args.warmup):
            ta.train(stream=sh, want_stats=False)
        m3, _, la, rays3 = timed(ta, k2, True)
        records["adaptive_controller"] = {"workload": WORKLOADS[args.workload].replace("4096 rays/step pinned", "rays/step set by the controller (2^18 compacted samples per step and GPU)"),
                                          "value": rays3 / (m3 * 1e-3), "unit": UNIT, "ms_per_step": m3 / k2, "steps": k2, "rays_per_step_last": int(la.n_rays), "samples_per_step_last": int(la.n_samples),
                                          "compacted_per_step_last": int(la.n_samples_compacted)}
        # (3) N > 1: BASELINE configs[3] — 16,384 rays/step in total, strong-sharded over the GPUs, global 2^18-sample budget as in the reference
        if world > 1:
            tsx, _ = make_testbed(16384, 1 << 18)
            for _ in range(args.pretrain + args.warmup):
                tsx.train(stream=sh, want_stats=False)
            m4, _, _, _ = timed(tsx, k2, False)
            records["strong_16384_rays"] = {"workload": WORKLOADS[args.workload].replace("4096 rays/step pinned", "16384 rays/step in total, ray-sharded over %d GPUs" % n_gpus), "scaling": "strong",
                                            "value": 16384 * k2 / (m4 * 1e-3), "unit": UNIT, "ms_per_step": m4 / k2, "steps": k2}
            # (4) the headline workload with the per-rank clamp / truncation / roll-over of round 1 (RNB_DP_EXACT=0, read when the communicator is installed):
            # what the two prefix all-gathers of the default — ONE sample order over all ranks, the single-GPU batch — cost
            prev = os.environ.get("RNB_DP_EXACT")
            os.environ["RNB_DP_EXACT"] = "0"
            try:
                tpr, _ = make_testbed(R, (1 << 18) * n_gpus)
            finally:
                if prev is None:
                    os.environ.pop("RNB_DP_EXACT", None)
                else:
                    os.environ["RNB_DP_EXACT"] = prev
            try:
                for _ in range(args.pretrain + args.warmup):
                    tpr.train(stream=sh, want_stats=False)
                m6, _, _, _ = timed(tpr, k2, False)
                records["per_rank_rule"] = {"workload": WORKLOADS[args.workload], "value": R * k2 / (m6 * 1e-3), "unit": UNIT, "ms_per_step": m6 / k2, "steps": k2, "nccl": tpr.comm_info(),
                                            "note": "RNB_DP_EXACT=0: every rank clamps / truncates / pads its own shard against target / world (not the single-GPU batch)"}
            except Exception as ex:      # noqa: BLE001  (a side record must not cost the headline line)
                records["per_rank_rule"] = {"error": str(ex)[:200]}
        line["records"] = records
    sys.stdout.flush(); os.dup2(_saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line)); sys.stdout.flush()
    if dist:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
