#!/usr/bin/env python
"""N-GPU check of the sharded optimizer protocol against the replicated (all-reduce) one on the CUDA library.  Launch under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_sharded_check.py
Both protocols train the same scene from the same initial parameters for K steps (float atomics make single steps differ in the last bits,
so the comparison is a tolerance, with a second replicated run as the yardstick), and after the all-gather every rank must hold
bit-identical binary16 parameters."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import rnb_loader, bench

K = int(os.environ.get("RNB_CHECK_STEPS", "40"))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = rnb_loader.load_package()
views, _ = bench.build_views(12, 320, 240, False)


class _Arr:
    def __init__(self, p, n, ts="<f4"): self.__cuda_array_interface__ = {"shape": (n,), "typestr": ts, "data": (p, False), "version": 3}


def run(mode):
    cfg = pkg.default_config(rays_per_batch=1024 * world, pin_rays_per_batch=1, world_size=world, rank=rank, target_batch_size=(1 << 16) * world)
    t = pkg.Testbed(cfg, pkg.default_flags(no_albedo=1))
    t.init_params(); t.load_training_data(views)
    gp, gn = t.grad_buffer(); sp, sn = t.stat_buffer()
    stat_t = torch.as_tensor(_Arr(sp, sn), device="cuda")
    pp, _, n, npad = t.param_buffers()
    par_t = torch.as_tensor(_Arr(pp, npad, "<f2"), device="cuda")
    if mode == "sharded":
        shard = npad // world
        grad_t = torch.as_tensor(_Arr(gp, npad), device="cuda")
        red_t = torch.zeros(shard, dtype=torch.float32, device="cuda"); own_t = torch.zeros(shard, dtype=torch.float16, device="cuda")
        t.set_optimizer_shard(rank * shard, (rank + 1) * shard, red_t.data_ptr())
    else:
        grad_t = torch.as_tensor(_Arr(gp, gn), device="cuda")
    losses = []
    for _ in range(K):
        ts = t.get_train_state()[0]
        if ts % min(max(ts // 16, 1), 16) == 0:
            t.training_prep_nerf()
        t.train_step_begin()
        if mode == "sharded":
            dist.reduce_scatter_tensor(red_t, grad_t); dist.all_reduce(stat_t)
            st = t.train_step_end()
            own_t.copy_(par_t[rank * shard:(rank + 1) * shard]); dist.all_gather_into_tensor(par_t, own_t)
        else:
            dist.all_reduce(grad_t); dist.all_reduce(stat_t)
            st = t.train_step_end()
        losses.append(float(st.loss))
    torch.cuda.synchronize()
    p = par_t[:n].clone()
    g = torch.as_tensor(_Arr(gp, npad), device="cuda").abs().max().item()      # the gradient buffer must be clean after the step
    ref = p.clone(); dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, p))
    flag = torch.tensor([1 if same else 0], device="cuda"); dist.all_reduce(flag)
    return p.float(), losses, int(flag.item()) == world, g


pa, la, sa, ga = run("allreduce")
pb, lb, sb, gb = run("allreduce")
pc, lc, sc, gc = run("sharded")
rel = lambda x, y: float((x - y).norm() / y.norm())
out = {"world": world, "steps": K, "run_to_run_allreduce": rel(pb, pa), "sharded_vs_allreduce": rel(pc, pa), "ranks_identical": [sa, sb, sc],
       "grad_buffer_abs_max_after": [ga, gb, gc], "loss_first_last": [[l[0], l[-1]] for l in (la, lb, lc)], "moved": rel(pa, torch.zeros_like(pa) + pa.mean()) > 0}
ok = sa and sb and sc and gc == 0.0 and out["sharded_vs_allreduce"] <= max(5 * out["run_to_run_allreduce"], 2e-3) and abs(lc[-1] - la[-1]) <= 0.05 * abs(la[-1]) + 1e-6
out["ok"] = bool(ok)
if rank == 0:
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dp_sharded_check.json"), "w"), indent=1)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
