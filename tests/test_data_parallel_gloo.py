"""Data-parallel step on CPU: two processes, gloo backend, the oracle's restatement of the ray-sharded step.

Rank r of G marches rays i == r (mod G) of the GLOBAL batch (same pcg32 streams, same image_idx as the single-GPU batch),
keeps every normalisation constant global (loss scale 128 / R_global, Eikonal divisor N_B), exchanges ONE sum all-reduce of
the gradient buffer (+ 4 statistics) between backward and Adam — exactly what bench.py does with NCCL around
rnb_train_step_begin / rnb_train_step_end.  Checked here:
  * the union of the two shards is the single-process batch (same rays, same samples),
  * the all-reduced gradient equals the sum of the two shard gradients computed in one process,
  * after the optimizer both ranks hold bit-identical parameters,
  * with no roll-over padding in play (n_in == target on both sides is not reachable on a toy scene, so the comparison is
    done on the un-padded weight-1 part: hash-grid touch set and MLP gradient direction agree with the single-process step).
"""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make(rank, world, threads=2):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rnb_loader
    from oracle_binding import Oracle, default_flags
    from common import SMALL
    scene = rnb_loader.load_scene()
    views = scene.make_scene(4, 64, 64, with_albedo=True)
    o = Oracle(threads=threads, **SMALL)
    o.init_params(1337, None)
    o.set_flags(default_flags(no_albedo=0, light_mode=-2)); o.set_views(views)
    o.set_world(world, rank)
    o.set_train_state(training_step=0, rays_per_batch=128, pin_rays=1)
    return o


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = _make(rank, world)
    out = {}
    for step in range(2):
        o.train_step_begin()
        g_local = o.get_grads().copy(); s_local = o.get_sums().copy()
        g = torch.from_numpy(g_local.copy()); s = torch.from_numpy(s_local.copy())
        dist.all_reduce(g); dist.all_reduce(s)                     # the single exchange of the step
        o.set_grads(g.numpy()); o.set_sums(s.numpy())
        st = o.train_step_end()
        if step == 0:
            out["g_local"] = g_local; out["g_sum"] = g.numpy().copy(); out["sums"] = s.numpy().copy(); out["loss"] = float(st.loss)
            out["ray_indices"] = o.last_losses()[0].copy()
    p = o.get_params()[0]
    gathered = [torch.zeros(p.size, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(p.copy()))
    out["params_equal"] = bool(all(torch.equal(gathered[0], x) for x in gathered))
    out["params"] = p.copy()
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=500) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # (1) shards partition the batch
    r0, r1 = res[0]["ray_indices"], res[1]["ray_indices"]
    assert np.all(r0 % 2 == 0) and np.all(r1 % 2 == 1)
    single = _make(0, 1)
    single.train_step_begin()
    rs = single.last_losses()[0]
    assert np.array_equal(np.sort(np.concatenate([r0, r1])), np.sort(rs)), "union of the shards != single-process batch"
    # (2) all-reduce == sum of the shard gradients, identical on both ranks
    assert np.array_equal(res[0]["g_sum"], res[1]["g_sum"])
    assert np.allclose(res[0]["g_sum"], res[0]["g_local"] + res[1]["g_local"], rtol=0, atol=0)
    assert res[0]["loss"] == res[1]["loss"]
    # (3) replicas stay bit-identical through the optimizer
    assert res[0]["params_equal"] and res[1]["params_equal"]
    assert np.array_equal(res[0]["params"], res[1]["params"])
    # (4) same touched hash entries and the same gradient direction as the single-process step (roll-over multiplicities
    #     are per rank, DESIGN.md §8, so magnitudes differ by the padding weights only)
    gs = single.get_grads(); gd = res[0]["g_sum"]
    grid = slice(single.off_grid, single.off_var)
    assert np.array_equal(gs[grid] != 0, gd[grid] != 0)
    mlp = slice(0, single.off_grid)
    cosv = float(gs[mlp].astype(np.float64) @ gd[mlp].astype(np.float64) / (np.linalg.norm(gs[mlp]) * np.linalg.norm(gd[mlp])))
    assert cosv > 0.999, cosv
    assert abs(float(res[0]["sums"][3]) - float(single.get_sums()[3])) < 1e-6     # global compacted count


# ---- sharded optimizer: reduce-scatter + Adam on 1/G of the parameters + all-gather of the binary16 weights ----------
def _worker_sharded(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rep = _make(rank, world, threads=1)          # replicated protocol: all-reduce, Adam everywhere (one thread: the oracle's float atomics
    shd = _make(rank, world, threads=1)          # sharded protocol                                  arrive in a fixed order, runs are bit-reproducible)
    n = rep.n_params
    npad = (n + 511) // 512 * 512; shard = npad // world
    b, e = rank * shard, min((rank + 1) * shard, n)
    shd.set_opt_shard(rank * shard, (rank + 1) * shard)
    out = {"equal_half": [], "equal_loss": []}
    for step in range(2):
        rep.train_step_begin()
        g = torch.from_numpy(rep.get_grads().copy()); s = torch.from_numpy(rep.get_sums().copy())
        dist.all_reduce(g); dist.all_reduce(s)
        rep.set_grads(g.numpy()); rep.set_sums(s.numpy())
        st_r = rep.train_step_end()

        shd.train_step_begin()
        gl = np.zeros(npad, np.float32); gl[:n] = shd.get_grads()
        g2 = torch.from_numpy(gl); s2 = torch.from_numpy(shd.get_sums().copy())
        dist.all_reduce(g2); dist.all_reduce(s2)          # gloo has no reduce-scatter: the shard of the all-reduced sum is the same numbers
        mine = np.zeros(n, np.float32); mine[b:e] = g2.numpy()[b:e]          # everything outside the shard is dropped, as in the kernel
        shd.set_grads(mine); shd.set_sums(s2.numpy())
        st_s = shd.train_step_end()
        hv = np.zeros(npad, np.float32); hv[:n] = shd.get_params()[1]
        own = torch.from_numpy(hv[rank * shard:(rank + 1) * shard].copy())
        parts = [torch.zeros(shard, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(parts, own)
        shd.set_half_params(torch.cat(parts).numpy()[:n])
        out["equal_half"].append(bool(np.array_equal(rep.get_params()[1], shd.get_params()[1])))
        out["equal_loss"].append(float(st_r.loss) == float(st_s.loss))
    m_r, _, e_r = rep.get_params(); m_s, _, e_s = shd.get_params()
    out["own_master_equal"] = bool(np.array_equal(m_r[b:e], m_s[b:e])); out["own_ema_equal"] = bool(np.array_equal(e_r[b:e], e_s[b:e]))
    out["dbg"] = (int((m_r[b:e] != m_s[b:e]).sum()), int((e_r[b:e] != e_s[b:e]).sum()), b, e, n, int(np.flatnonzero(m_r[b:e] != m_s[b:e])[:1].sum()) if (m_r[b:e] != m_s[b:e]).any() else -1, int(np.isnan(m_r).sum()), int(np.isnan(e_r).sum()))
    other = np.ones(n, bool); other[b:e] = False
    out["other_master_untouched"] = bool(np.array_equal(m_s[other], _make(rank, world).get_params()[0][other]))
    out["changed"] = bool(np.any(m_r != _make(rank, world).get_params()[0]))
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_sharded_optimizer_matches_replicated():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=800) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert all(res[r]["equal_half"]), "training weights differ between the replicated and the sharded protocol"
        assert all(res[r]["equal_loss"])
        assert res[r]["own_master_equal"] and res[r]["own_ema_equal"], res[r].get("dbg")
        assert res[r]["other_master_untouched"] and res[r]["changed"]


# ---- one sample order over all ranks (the product's default with its own communicator): the shards ARE the single-process batch ------
def test_one_sample_order_reproduces_the_single_process_step():
    """orc_set_dp_exact: clamp, 2^18 truncation and roll-over multiplicity are taken at the sample's index in the batch of all ranks.  The sum of the
    shard gradients is then the single-process gradient (up to the order of the float additions) and the trajectories coincide; with the per-rank rule
    the multiplicities differ and the trajectories drift (the round-1 approximation, kept as RNB_DP_EXACT=0)."""
    def make(rank, world, exact):
        o = _make(rank, world, threads=2)
        o.set_dp_exact(exact)
        o.set_train_state(training_step=0, rays_per_batch=64, pin_rays=1)
        return o

    def run(world, exact, steps):
        ranks = [make(r, world, exact) for r in range(world)]
        g0 = None
        for _ in range(steps):
            for o in ranks:
                o.train_step_begin()
            g = np.sum([o.get_grads().astype(np.float64) for o in ranks], axis=0).astype(np.float32)
            sm = np.sum([o.get_sums() for o in ranks], axis=0)
            if g0 is None:
                g0 = g.copy()
            for o in ranks:
                o.set_grads(g); o.set_sums(sm); st = o.train_step_end()
        return ranks[0].get_params()[0].copy(), g0, st

    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
    p1, g1, s1 = run(1, 0, 2)
    pe, ge, se = run(2, 1, 2)
    pa, ga, sa = run(2, 0, 2)
    assert rel(ge, g1) < 1e-5, rel(ge, g1)                # same samples, same multiplicities
    assert rel(pe, p1) < 1e-5, rel(pe, p1)
    assert abs(se.loss - s1.loss) < 1e-6 * max(1.0, abs(s1.loss))
    assert int(se.n_compacted) == int(s1.n_compacted) and int(se.n_samples) == int(s1.n_samples)
    assert rel(ga, g1) > 10 * rel(ge, g1)                 # the per-rank rule is an approximation (and measurably so)
    assert np.isfinite(pa).all() and rel(pa, p1) < 0.1


# ---- the prefix-table exchange behind "one sample order" (rnb_common.cuh: foreign_prefix / global_total; rnb_march.cu: k_prefix_positions) ----------
def _prefix_worker(rank, world, port, n_rays, seed, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = np.random.default_rng(seed).integers(0, 1025, size=n_rays).astype(np.int64)      # per-ray sample counts of the GLOBAL batch (0 = ray misses the grid)
    L = (n_rays + world - 1) // world
    mine = np.zeros(L, np.int64)                                                              # position m of this rank = global ray m * world + rank
    own = counts[rank::world]; mine[:own.size] = own
    incl = np.cumsum(mine)                                                                    # k_prefix_positions
    table = [torch.zeros(L, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(table, torch.from_numpy(incl))                                            # ncclAllGather in the product
    xg = np.stack([t.numpy() for t in table])
    base = np.zeros(L, np.int64)
    for m in range(L):                                                                        # foreign_prefix + the rank's own exclusive prefix
        s = sum(xg[r][m] for r in range(rank))
        if m:
            s += sum(xg[r][m - 1] for r in range(rank + 1, world))
        base[m] = s + incl[m] - mine[m]
    total = int(sum(xg[r][L - 1] for r in range(world)))                                      # global_total
    q.put((rank, base[:own.size].copy(), total))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rays", [(2, 257), (3, 1000), (3, 2)])
def test_prefix_table_exchange_gives_every_ray_its_place_in_the_global_batch(world, n_rays):
    """Interleaved rays, one all-gather of per-position inclusive prefixes: rank r's ray at position m starts at the exclusive prefix of the GLOBAL ray order
    (what one GPU's ordered scan hands out) — for any world size, with ragged shards (n_rays % world != 0) and with fewer rays than ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 7 * world + n_rays) % 2000)
    procs = [ctx.Process(target=_prefix_worker, args=(r, world, port, n_rays, 1234, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, base, total = q.get(timeout=120)
        res[r] = (base, total)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    counts = np.random.default_rng(1234).integers(0, 1025, size=n_rays).astype(np.int64)
    excl = np.cumsum(counts) - counts
    for r in range(world):
        assert np.array_equal(res[r][0], excl[r::world]), r
        assert res[r][1] == int(counts.sum())
