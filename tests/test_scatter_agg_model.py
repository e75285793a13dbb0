"""Lane-level model of the warp-aggregated hash scatter (csrc/rnb_encode.cuh: scatter_level_agg): runs of adjacent lanes with the same
cell key are reduced with a segmented shuffle-down scan and only the first lane of a run issues the reductions.  The model executes the
kernel's exact control flow (shfl.up of the key, ballot of run heads, ffs / redux.max for the run bounds, log2(run) shfl.down steps with
the `lane + d <= seg_last` guard) on numpy lanes and checks it against a per-key sum: what reaches memory must be unchanged."""
import numpy as np
import pytest


def ffs(x):
    return (x & -x).bit_length()


def aggregate(keys, vals, live):
    """returns {(lane, key): summed value} for the lanes that issue reductions, following scatter_level_agg step by step"""
    keys = np.where(live, keys, 0xFFFFFFFF).astype(np.uint64)
    v = np.where(live[:, None], vals, 0.0).astype(np.float64).copy()
    lane = np.arange(32)
    prev = np.concatenate([[keys[0]], keys[:-1]])                    # __shfl_up_sync(.., 1): lane 0 reads its own value
    head = (lane == 0) | (prev != keys)
    heads = int(sum(1 << int(i) for i in lane[head]))                # __ballot_sync
    issued = {}
    if 32 - bin(heads).count("1") >= 4:
        seg_last = np.zeros(32, np.int64)
        for i in range(32):
            above = 0 if i == 31 else heads >> (i + 1)
            seg_last[i] = i + ffs(above) - 1 if above else 31
        run = int(np.max(seg_last - lane))                           # __reduce_max_sync
        d = 1
        while d <= run:
            other = np.concatenate([v[d:], v[:d]])                   # __shfl_down_sync: out-of-range lanes return their own value
            other[32 - d:] = v[32 - d:]
            take = lane + d <= seg_last
            v[take] += other[take]
            d <<= 1
        for i in lane[head]:
            issued[(int(i), int(keys[i]))] = v[i]
    else:
        for i in lane:
            issued[(int(i), int(keys[i]))] = v[i]
    return issued


@pytest.mark.parametrize("seed", range(40))
def test_segmented_reduction_preserves_per_key_sums(seed):
    rng = np.random.default_rng(seed)
    runs = []
    mean_run = [1, 2, 5, 12, 40][seed % 5]
    while sum(runs) < 32:
        runs.append(int(rng.integers(1, 2 * mean_run + 1)))
    keys = np.repeat(rng.integers(0, 6, len(runs)), runs)[:32]      # equal keys may also recur in separate runs: each run issues its own sum
    vals = rng.standard_normal((32, 16))
    live = rng.random(32) > (0.0 if seed % 3 else 0.2)
    issued = aggregate(keys, vals, live)
    # (1) what reaches memory per key is the plain per-key sum over live lanes
    for k in set(int(x) for x in keys[live]):
        got = sum(val for (ln, kk), val in issued.items() if kk == k)
        want = vals[live & (keys == k)].sum(axis=0)
        assert np.allclose(got, want, atol=1e-12)
    # (2) dead lanes contribute nothing under their private key
    dead = [val for (ln, kk), val in issued.items() if kk == 0xFFFFFFFF]
    assert all(np.all(d == 0.0) for d in dead)
    # (3) issuing lanes are run heads: never two adjacent issuing lanes with the same key when the reduction ran
    if len(issued) < 32:
        lanes = sorted(ln for ln, _ in issued)
        kk = np.where(live, keys, 0xFFFFFFFF)
        assert all(kk[a] != kk[a - 1] for a in lanes if a > 0)


def test_threshold_keeps_the_plain_path_for_mostly_distinct_cells():
    keys = np.arange(32); keys[5] = keys[4]; keys[20] = keys[19]     # two pairs only: 2 lanes saved < 4
    vals = np.ones((32, 16)); live = np.ones(32, bool)
    issued = aggregate(keys, vals, live)
    assert len(issued) == 32


def test_whole_warp_in_one_cell_issues_once():
    keys = np.full(32, 7); vals = np.arange(32 * 16, dtype=np.float64).reshape(32, 16); live = np.ones(32, bool)
    issued = aggregate(keys, vals, live)
    assert list(issued) == [(0, 7)] and np.allclose(issued[(0, 7)], vals.sum(axis=0))
