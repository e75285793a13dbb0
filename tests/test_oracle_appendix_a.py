"""SURVEY.md Appendix A, rule by rule, as known-answer tests of the oracle: each rule is restated here INDEPENDENTLY (numpy, from the text of the appendix
and the reference lines it cites) and compared with what oracle/rnb_oracle.cpp computes.  The hash tables, pcg32, Morton codes and the level schedule
have their own file (test_oracle_kat.py); the derivatives are checked by finite differences (test_oracle_gradcheck.py); the reference build itself pins
the rest (test_reference_golden.py).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from common import SMALL
from oracle_binding import Oracle, default_flags, lib


def _next_multiple(v, d):
    return (v + d - 1) // d * d


# ---- roll-over (ref:dependencies/neus2_tcnn/include/tiny-cuda-nn/common_device.h:514-535; src/testbed_nerf.cu:4044-4052) ----------------------------------
@pytest.mark.parametrize("n_in,n_batch", [(1, 8), (3, 8), (5, 16), (8, 8), (9, 8), (100, 256), (255, 256), (37, 1000)])
def test_rollover_weight_is_the_multiplicity_of_the_cyclic_padding(n_in, n_batch):
    """The batch is padded to n_batch with cyclic copies of the first n_in samples; the ORIGINAL keeps its gradient, every COPY gets it times n_in / n_batch.
    Summing over the copies of one source sample gives the weight the oracle (and the product's backward) applies to that sample once."""
    L = lib()
    w = np.zeros(min(n_in, n_batch), np.float64)
    for i in range(n_batch):
        if n_in >= n_batch:
            if i < len(w):
                w[i] += 1.0
            continue
        w[i % n_in] += 1.0 if i < n_in else float(np.float32(n_in) / np.float32(n_batch))
    got = np.array([L.orc_rollover_weight(C.c_uint32(s), C.c_uint32(n_in), C.c_uint32(n_batch)) for s in range(len(w))], np.float64)
    assert np.allclose(got, w, rtol=2e-6, atol=0), (got, w)


# ---- optimizer (ref:configs/nerf/base.json:5-29; tcnn optimizers/adam.h:88-199, ema.h:116-152, exponential_decay.h:61-72) ----------------------------------
def _half(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def test_adam_sparse_rule_ema_debias_and_lr_decay():
    o = Oracle(threads=1, **SMALL)
    o.init_params(7, None)
    n, n_mlp, off_var = o.n_params, o.off_grid, o.off_var
    master0, _, ema0 = o.get_params()
    rs = np.random.RandomState(3)
    lr, b1, b2, eps, l2, scale, decay = 1e-3, 0.9, 0.99, 1e-15, 1e-6, 128.0, 0.95
    w = master0.astype(np.float32).copy(); m1 = np.zeros(n, np.float32); m2 = np.zeros(n, np.float32); cnt = np.zeros(n, np.uint32); ema = ema0.copy()
    touched_once = None
    for step in range(1, 4):
        g = np.zeros(n, np.float32)
        g[:n_mlp] = rs.normal(0, 3.0, n_mlp)
        g[:7] = 0.0                                                     # MLP weights with a zero gradient still decay through L2
        idx = rs.choice(np.arange(n_mlp, off_var), 500, replace=False)  # a sparse set of hash entries; the others are skipped entirely
        g[idx] = rs.normal(0, 50.0, idx.size)
        if step == 1:
            touched_once = idx[:50]
        else:
            g[touched_once] = 0.0                                       # touched in step 1 only: no moment decay, no step count afterwards
        g[off_var] = 0.25 * step
        o.set_grads(g); o.optimizer_step()
        # restatement
        grad = _half(g) / np.float32(scale)
        is_mlp = np.arange(n) < n_mlp
        active = is_mlp | (grad != 0)
        grad = np.where(is_mlp, grad + np.float32(l2) * w, grad).astype(np.float32)
        one = np.float32(1)      # the reference forms 1 - beta in float (adam.h:150-160): 1 - 0.9f is not float(0.1)
        m1n = (np.float32(b1) * m1 + (one - np.float32(b1)) * grad).astype(np.float32)
        m2n = (np.float32(b2) * m2 + (one - np.float32(b2)) * (grad * grad)).astype(np.float32)
        cn = cnt + 1
        corr = (np.sqrt(1 - np.power(np.float32(b2), cn.astype(np.float32))) / (1 - np.power(np.float32(b1), cn.astype(np.float32)))).astype(np.float32)
        eff = (np.float32(lr) * corr / (np.sqrt(m2n) + np.float32(eps))).astype(np.float32)
        wn = (w - eff * m1n).astype(np.float32)
        w = np.where(active, wn, w); m1 = np.where(active, m1n, m1); m2 = np.where(active, m2n, m2); cnt = np.where(active, cn, cnt).astype(np.uint32)
        d_old = np.float32(1 - decay ** (step - 1)); d_new = np.float32(1.0 / (1 - decay ** step))
        ema = _half((ema * np.float32(decay) * d_old + _half(w) * (np.float32(1) - np.float32(decay))) * d_new)
    master, half, ema_got = o.get_params(); gm1, gm2, gcnt = o.get_opt_state()
    assert np.array_equal(gcnt, cnt)
    assert int(cnt[touched_once].max()) == 1 and int(cnt[:n_mlp].min()) == 3 and int(cnt[off_var]) == 3
    untouched = np.setdiff1d(np.arange(n_mlp, off_var), np.nonzero(cnt)[0])
    assert untouched.size > 0 and np.array_equal(master[untouched], master0[untouched])            # never touched: bit-identical weights
    assert np.allclose(gm1, m1, rtol=2e-6, atol=0) and np.allclose(gm2, m2, rtol=2e-6, atol=0)      # bit-equal on this platform; two ulps for another libm's powf
    assert np.allclose(master, w, rtol=2e-6, atol=0)
    assert np.array_equal(half, _half(master))
    assert np.allclose(ema_got, ema, rtol=2e-3, atol=1e-7)                                        # binary16 storage: one ulp of slack
    assert not np.array_equal(master[:7], master0[:7])                                             # L2 moved the zero-gradient MLP weights


# ---- controller + refresh cadence (ref:src/testbed_nerf.cu:3554-3555,3906-3911,4125-4138; src/testbed.cu:2805-2806) ----------------------------------
def test_controller_formula_and_refresh_cadence():
    import rnb_loader
    scene = rnb_loader.load_scene()
    views = scene.make_scene(4, 64, 64, with_albedo=False)
    o = Oracle(threads=2, **SMALL)
    o.init_params(1337, None)
    o.set_flags(default_flags(no_albedo=1, light_mode=-2)); o.set_views(views)
    target = 1 << 12
    o.set_train_state(training_step=0, rays_per_batch=256, pin_rays=0, target_batch=target)
    R = 256
    for _ in range(3):
        st = o.train_step()
        assert st.n_compacted > 0
        want = min(_next_multiple(int(np.float32(R) * np.float32(target) / np.float32(st.n_compacted)), 128), 1 << 18)
        assert st.rays_per_batch_next == want, (st.rays_per_batch_next, want, R, st.n_compacted)
        R = st.rays_per_batch_next
    # occupancy refresh: every clamp(step / 16, 1, 16) steps of the canonical training step
    L = lib()
    due = []
    for step in (16, 32, 33, 34, 47, 256, 257, 271, 272, 4095):      # every due step costs one refresh of the 128^3 grid
        o.set_train_state(training_step=step, rays_per_batch=128, pin_rays=1, target_batch=target)
        o.set_canonical_state(step, 4)
        due.append((step, int(L.orc_prep_if_due(o.h))))
    want = [(s, 1 if s % min(max(s // 16, 1), 16) == 0 else 0) for s, _ in due]
    assert due == want
    assert dict(due)[32] == 1 and dict(due)[33] == 0 and dict(due)[34] == 1 and dict(due)[257] == 0 and dict(due)[272] == 1


# ---- density of the occupancy grid and the logistic variance (ref:src/common_operation.cuh:310-328; nerf_network.h:70,689-694) ----------------------------------
def test_grid_density_formula_in_binary16():
    o = Oracle(threads=1, **SMALL)
    o.init_params(11, None)
    xyz = np.random.RandomState(5).uniform(0.2, 0.8, (256, 3)).astype(np.float32)
    vl = o.valid_level(0)
    sdf, dens = o.eval_sdf(xyz, vl)
    _, half, _ = o.get_params()
    var = np.float16(half[o.off_var])
    assert float(var) == pytest.approx(0.3, abs=1e-3)                                              # init 0.3 (binary16)
    h = np.float16
    s = h(np.exp(np.float32(h(var * h(10.0)))))                                                    # inv_s = exp(10 var), the product in binary16
    sg = (1.0 / (1.0 + np.exp(-np.float32(h(sdf.astype(np.float16) * s))))).astype(np.float16)     # logistic of the binary16 product
    want = ((s * sg).astype(np.float16) * (h(1.0) - sg).astype(np.float16)).astype(np.float32)
    assert np.allclose(dens, want, rtol=2e-3, atol=1e-6), float(np.max(np.abs(dens - want)))
    assert float(dens.max()) <= float(s) / 4 * 1.01                                                # s sigma (1 - sigma) <= s / 4
