"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): occupancy / sample indices bit-exact; floating-point outputs within 1e-3 relative.
Outputs that are stored in binary16 (the 16-wide network output, dL/doutput) are compared with a norm-wise 1e-3 bound
plus an element-wise bound of two binary16 ulps, because one rounding flip of a binary16 value is already 9.8e-4.
"""
import numpy as np
import pytest
from common import SMALL, MID, FULL, make_pair, rel_err
from oracle_binding import default_flags as orc_flags

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star tolerance for fp32-relative comparisons


def half_close(a, b, ulps=2.0, atol=1e-6):
    """element-wise: |a-b| <= ulps * ulp_fp16(max(|a|,|b|)) (+atol)"""
    m = np.maximum(np.abs(a), np.abs(b))
    ulp = np.where(m > 0, 2.0 ** (np.floor(np.log2(np.maximum(m, 2.0 ** -14))) - 10), 2.0 ** -24)
    return np.abs(a - b) <= ulps * ulp + atol


@pytest.fixture(scope="module")
def small_scene(scene_mod):
    return scene_mod.make_scene(6, 96, 96, with_albedo=True)


def _occupancy_from_oracle(o, t):
    """One density-grid refresh on the oracle; the product gets the oracle's bitfield so that marching is compared on identical occupancy."""
    o.set_train_state(training_step=0, rays_per_batch=512)
    o.density_update(128 ** 3, 0, o.valid_level(0))
    bf = o.get_bitfield()
    t.set_bitfield(bf)
    return bf


@pytest.mark.parametrize("cfg", [SMALL, MID])
def test_march_samples_bit_exact(pkg, small_scene, cfg):
    o, t = make_pair(pkg, cfg, views=small_scene)
    bf = _occupancy_from_oracle(o, t)
    assert 0 < np.unpackbits(bf[:128 ** 3 // 8]).sum() < 128 ** 3
    for n_rays, nrt, max_samples in [(512, 0, 1 << 20), (512, 512, 30000), (300, 77, 1 << 20)]:
        a = o.generate_samples(n_rays, nrt, max_samples)
        b = t.stage_generate(n_rays, nrt, max_samples)
        assert a["n_kept"] == b["n_kept"] and a["n_samples"] == b["n_samples"] and a["n_kept"] > 0
        assert np.array_equal(a["ray_indices"], b["ray_indices"])
        assert np.array_equal(a["numsteps"], b["numsteps"])
        n = int((a["numsteps"][:, 0] + a["numsteps"][:, 1]).max())
        # positions and directions bit-exact => occupancy cell indices bit-exact
        assert np.array_equal(a["coords"][:n].view(np.uint32), b["coords"][:n].view(np.uint32))
        assert np.array_equal(a["rays"].view(np.uint32), b["rays"].view(np.uint32))


def test_march_empty_and_full_occupancy(pkg, small_scene):
    o, t = make_pair(pkg, SMALL, views=small_scene)
    for fill in (0x00, 0xFF):
        bf = np.full(128 ** 3, fill, np.uint8)
        o.set_bitfield(bf); t.set_bitfield(bf)
        a = o.generate_samples(256, 0, 1 << 20); b = t.stage_generate(256, 0, 1 << 20)
        assert a["n_kept"] == b["n_kept"] and a["n_samples"] == b["n_samples"]
        if fill == 0:
            assert a["n_kept"] == 0
        else:
            assert a["numsteps"][:, 0].max() <= 1024
            assert np.array_equal(a["numsteps"], b["numsteps"])
            n = int((a["numsteps"][:, 0] + a["numsteps"][:, 1]).max())
            assert np.array_equal(a["coords"][:n].view(np.uint32), b["coords"][:n].view(np.uint32))


# FULL = the shipped default network (L=14, T=2^19: levels 0-4 dense incl. the non-power-of-two resolutions 34 / 50 / 72, levels 5-13 hashed);
# steps 0 / 300 / 700 have 14 / 7 / 14 live levels (grid.h:1430-1437)
@pytest.mark.parametrize("cfg,step", [(SMALL, 0), (MID, 0), (MID, 300), (FULL, 0), (FULL, 300), (FULL, 700)])
def test_network_forward(pkg, cfg, step):
    o, t = make_pair(pkg, cfg, seed_params=1)
    t.set_train_state(step, 512); vl = o.valid_level(step)
    rs = np.random.RandomState(5)
    coords = rs.rand(3000, 7).astype(np.float32)
    coords[0, :3] = 0.0; coords[1, :3] = 1.0; coords[2, :3] = 0.5          # cube corners / centre
    ref, nref = o.network_forward(coords, vl)
    out, nrm = t.stage_forward(coords)
    assert rel_err(nrm, nref) < TOL
    for cols in ([0, 1, 2], [3], [4, 5, 6], [7], [8, 9, 10]):
        assert rel_err(out[:, cols], ref[:, cols]) < TOL, cols
    assert half_close(out[:, :11], ref[:, :11], ulps=4).mean() > 0.999
    assert np.array_equal(out[:, 7:11], ref[:, 7:11])                       # variance and view dir are pure copies


def test_network_forward_ema_weights_and_empty_batch(pkg):
    o, t = make_pair(pkg, SMALL, seed_params=2)
    out, _ = t.stage_forward(np.zeros((0, 7), np.float32))
    assert out.shape == (0, 16)
    # EMA copy is zero before the first optimizer step: sdf = bias, normals 0
    coords = np.random.RandomState(0).rand(64, 7).astype(np.float32)
    out, _ = t.stage_forward(coords, use_ema=True)
    ref, _ = o.network_forward(coords, o.valid_level(0), use_ema=True)
    assert np.allclose(out[:, :8], ref[:, :8], atol=1e-6)


@pytest.mark.parametrize("flagset", [dict(no_albedo=1), dict(no_albedo=0), dict(no_albedo=0, apply_L2=0), dict(no_albedo=0, apply_rgbplus=0, apply_relu=1),
                                     dict(no_albedo=1, apply_supernormal=1, light_opti=1), dict(no_albedo=0, apply_bce=1, light_mode=2)])
def test_loss_and_output_gradients(pkg, small_scene, flagset):
    f = orc_flags(**flagset)
    o, t = make_pair(pkg, SMALL, views=small_scene, flags=f, seed_params=3)
    _occupancy_from_oracle(o, t)
    g = o.generate_samples(384, 0, 1 << 20)
    K = g["n_kept"]; n = int((g["numsteps"][:, 0] + g["numsteps"][:, 1]).max())
    out_a, _ = o.network_forward(g["coords"][:n], o.valid_level(0))
    # make the transmittance cut bite: scale sdf down so that alphas are large on some rays
    out_a[:, 3] *= 0.2
    nf, cb, ne, total = o.compact(out_a, g["numsteps"], 20000)
    assert (nf < g["numsteps"][:, 0]).any() or total > 0
    oc = np.zeros((total, 16), np.float32)
    for k in range(K):
        oc[cb[k]:cb[k] + nf[k]] = out_a[g["numsteps"][k, 1]:g["numsteps"][k, 1] + nf[k]]
    d_ref, l_ref, e_ref, m_ref = o.loss(oc, g["ray_indices"], nf, cb, ne, 384, 0, 0)
    d, l, e, m = t.stage_loss(oc, g["ray_indices"], nf, cb, ne, 384, 0)
    assert rel_err(l, l_ref) < TOL and rel_err(e, e_ref) < TOL and rel_err(m, m_ref) < TOL
    emitted = np.zeros(total, bool)
    for k in range(K):
        emitted[cb[k]:cb[k] + ne[k]] = True
    assert emitted.any() and not emitted.all()          # truncation at max_compacted exercised
    for cols in ([0, 1, 2], [3], [4, 5, 6], [7], [8, 9, 10]):
        a, b = d[emitted][:, cols], d_ref[emitted][:, cols]
        # classic BCE divides by (1 - weight_sum): ill-conditioned near saturated rays, fp32 noise is amplified
        tol = 2e-2 if flagset.get("apply_bce") else 2 * TOL
        if np.linalg.norm(b) > 0:
            assert rel_err(a, b) < tol, (cols, rel_err(a, b))
    assert np.all(d[~emitted] == 0)


def ray_like_coords(rs, n_rays, per_ray):
    """consecutive lattice points of rays (dt = sqrt(3)/1024), the order compacted samples have in a training step: adjacent samples share
    coarse cells, which is what the warp-aggregated scatter and the paired 16-byte reductions of the backward rely on"""
    o = rs.uniform(0.2, 0.8, (n_rays, 1, 3)); d = rs.randn(n_rays, 1, 3); d /= np.linalg.norm(d, axis=2, keepdims=True)
    tt = (np.arange(per_ray) - per_ray / 2)[None, :, None] * (np.sqrt(3.0) / 1024)
    pos = np.clip(o + d * tt, 0.0, 1.0).reshape(-1, 3)
    c = rs.rand(n_rays * per_ray, 7).astype(np.float32)
    c[:, :3] = pos.astype(np.float32)
    return c


@pytest.mark.parametrize("cfg,step,raylike", [(SMALL, 0, False), (MID, 0, False), (MID, 200, False), (FULL, 0, False), (FULL, 300, True), (FULL, 700, True), (FULL, 700, False)])
def test_network_backward_first_and_second_order(pkg, cfg, step, raylike):
    o, t = make_pair(pkg, cfg, seed_params=4)
    t.set_train_state(step, 512); vl = o.valid_level(step)
    rs = np.random.RandomState(9)
    n = 4096
    coords = ray_like_coords(rs, 32, 128) if raylike else rs.rand(n, 7).astype(np.float32)
    dout = (rs.randn(n, 16) * 0.05).astype(np.float16).astype(np.float32)
    dout[:, 11:] = 0
    n_in = 3000                                          # roll-over multiplicities differ across the batch
    g_ref = o.network_backward(coords, dout, n_in, 1 << 18, vl)
    g = t.stage_backward(coords, dout, n_in)
    sl = dict(sdf=slice(o.off_sdf, o.off_rgb), rgb=slice(o.off_rgb, o.off_grid), grid=slice(o.off_grid, o.off_var), var=slice(o.off_var, o.off_var + 1))
    for name, s in sl.items():
        assert np.linalg.norm(g_ref[s]) > 0
        # gradients are sums of products of binary16-rounded activations: the tensor-core path accumulates in a different
        # order than the oracle, single binary16 rounding flips (9.8e-4 each) propagate through three layers -> 5e-3 norm-wise
        assert rel_err(g[s], g_ref[s]) < 5 * TOL, (name, rel_err(g[s], g_ref[s]))
    # same set of touched hash entries (index arithmetic is exact)
    assert np.array_equal(g[sl["grid"]] != 0, g_ref[sl["grid"]] != 0) or rel_err((g[sl["grid"]] != 0).astype(float), (g_ref[sl["grid"]] != 0).astype(float)) < 1e-3


@pytest.mark.parametrize("agg,pair", [(0, 0), (0, 1), (5, 0), (8, 1)])
def test_backward_scatter_variants_agree(pkg, agg, pair, monkeypatch):
    """The warp-aggregated scatter (RNB_SCATTER_AGG levels) and the paired 16-byte reductions (RNB_SCATTER_PAIR) are re-orderings of the
    same sums: every setting must give the default setting's gradients (fp32 accumulation order differs: 1e-5) on the default network with
    ray-ordered samples, and the same set of touched entries."""
    rs = np.random.RandomState(21)
    coords = ray_like_coords(rs, 48, 96)
    n = coords.shape[0]
    dout = (rs.randn(n, 16) * 0.05).astype(np.float16).astype(np.float32); dout[:, 11:] = 0
    res = []
    for env in (None, (agg, pair)):
        if env is None:
            monkeypatch.delenv("RNB_SCATTER_AGG", raising=False); monkeypatch.delenv("RNB_SCATTER_PAIR", raising=False)
        else:
            monkeypatch.setenv("RNB_SCATTER_AGG", str(env[0])); monkeypatch.setenv("RNB_SCATTER_PAIR", str(env[1]))
        o, t = make_pair(pkg, FULL, seed_params=4)
        t.set_train_state(700, 512)
        res.append(t.stage_backward(coords, dout, n - 500).copy())
        del t, o
    g0, g1 = res
    assert np.linalg.norm(g0) > 0
    assert rel_err(g1, g0) < 1e-5, rel_err(g1, g0)
    assert np.array_equal(g1 != 0, g0 != 0)


@pytest.mark.parametrize("cfg", [SMALL, FULL])
def test_optimizer_adam_ema_sparse_rule(pkg, cfg):
    o, t = make_pair(pkg, cfg, seed_params=6)
    rs = np.random.RandomState(1)
    for it in range(3):
        g = np.zeros(o.n_params, np.float32)
        g[:o.off_grid] = rs.randn(o.off_grid) * 0.5
        touched = rs.rand(o.off_var - o.off_grid) < 0.3
        g[o.off_grid:o.off_var] = np.where(touched, rs.randn(touched.size) * 1e-2, 0)
        g[o.off_grid + 5] = 1e-9                              # underflows in binary16 -> must be skipped
        g[o.off_var] = 0.7
        o.set_grads(g); o.optimizer_step()
        t.stage_optimizer(g)
        m_ref, h_ref, e_ref = o.get_params()
        p = t.get_params()
        assert np.allclose(p, m_ref, rtol=1e-5, atol=1e-7)
        assert rel_err(t.export_params_fp16(False).astype(np.float32), h_ref) < 1e-4
        assert rel_err(t.export_params_fp16(True).astype(np.float32), e_ref) < 1e-3
    untouched = ~touched
    untouched[5] = True
    p0 = t.get_params()
    # entries that never saw a (representable) gradient keep their initial value
    assert np.array_equal(p0[o.off_grid:o.off_var][5], m_ref[o.off_grid:o.off_var][5])


def test_density_grid_bitfield(pkg, small_scene):
    o, t = make_pair(pkg, SMALL, views=small_scene)
    o.set_train_state(training_step=0, rays_per_batch=512); t.set_train_state(0, 512)
    t.set_rng(o.get_rng())
    o.density_update(128 ** 3, 0, o.valid_level(0))
    t.training_prep_nerf()
    g_ref = o.get_density_grid(); g, ema_step = t.export_density_grid()
    assert ema_step == 1
    assert rel_err(g, g_ref) < TOL
    b_ref = np.unpackbits(o.get_bitfield()); b = np.unpackbits(t.get_bitfield())
    diff = int((b != b_ref).sum())
    # cell indices are exact; a bit may only differ where the density sits within rounding of the threshold
    assert diff <= 1e-4 * b.size, diff
    # mips are OR-pools of mip 0: every set bit in mip 1 has a set parent region
    assert b[128 ** 3:2 * 128 ** 3].sum() > 0


def test_full_training_steps_track_the_oracle(pkg, small_scene):
    f = orc_flags(no_albedo=0, light_mode=1)
    o, t = make_pair(pkg, SMALL, views=small_scene, flags=f, rays_per_batch=256)
    o.set_train_state(training_step=0, rays_per_batch=256, pin_rays=1); t.set_train_state(0, 256)
    t.set_rng(o.get_rng())
    for it in range(3):
        a = o.train_step()
        b = t.train()
        assert b.training_step == it + 1
        # identical occupancy => identical sample counts, unless a borderline density bit flipped
        assert abs(int(a.n_samples) - int(b.n_samples)) <= 0.01 * a.n_samples
        assert abs(int(a.n_compacted) - int(b.n_samples_compacted)) <= 0.01 * a.n_compacted
        assert abs(a.loss - b.loss) <= 0.02 * abs(a.loss) + 1e-6
        assert abs(a.mask_loss - b.mask_loss) <= 0.02 * abs(a.mask_loss) + 1e-6
    m_ref, _, _ = o.get_params()
    p = t.get_params()
    # three Adam steps move every touched weight by ~3e-3; trajectories agree to a fraction of that
    assert np.abs(p[:o.off_grid] - m_ref[:o.off_grid]).max() < 2e-3
    assert rel_err(p[:o.off_grid], m_ref[:o.off_grid]) < 5e-3


@pytest.mark.parametrize("use_ema", [False, True])
def test_eval_sdf_matches_oracle(pkg, use_ema):
    """rnb_eval_sdf (NerfNetwork::sdf / ::density, nerf_network.h:454-537) on device buffers: the tcgen05 probe path (sdf, density)
    and the CUDA-core path (with normals) against the oracle."""
    import torch
    o, t = make_pair(pkg, MID, seed_params=11)
    if use_ema:                               # snapshot hand-off: EMA weights := training weights (trainer.h:263-275)
        h = t.export_params_fp16(use_ema=False); t.import_params_fp16(h)
        m = np.asarray(h).view(np.float16).astype(np.float32); o.set_params(m)
    rs = np.random.RandomState(3)
    xyz = rs.uniform(0.05, 0.95, (40000, 3)).astype(np.float32)
    vl = o.valid_level(0)
    s_ref, d_ref = o.eval_sdf(xyz, vl, use_ema=False)
    x = torch.from_numpy(xyz).cuda()
    sdf = torch.empty(xyz.shape[0], device="cuda"); dens = torch.empty_like(sdf); nrm = torch.empty(xyz.shape[0], 3, device="cuda")
    t.eval_sdf_device(x.data_ptr(), xyz.shape[0], sdf.data_ptr(), None, dens.data_ptr(), use_ema=use_ema)
    torch.cuda.synchronize()
    assert rel_err(sdf.cpu().numpy(), s_ref) < TOL
    assert rel_err(dens.cpu().numpy(), d_ref) < 5e-3          # binary16 arithmetic of sdf_to_density_variance_buffer
    sdf2 = torch.empty_like(sdf)
    t.eval_sdf_device(x.data_ptr(), xyz.shape[0], sdf2.data_ptr(), nrm.data_ptr(), None, use_ema=use_ema)
    torch.cuda.synchronize()
    assert rel_err(sdf2.cpu().numpy(), s_ref) < TOL
    assert np.isfinite(nrm.cpu().numpy()).all()
    assert half_close(sdf2.cpu().numpy(), sdf.cpu().numpy(), ulps=2.0).mean() > 0.999      # the two kernels agree to binary16 rounding


def test_checkpoint_restore_replays_the_same_steps(pkg, small_scene):
    """rnb_checkpoint_save / _restore: the complete training state (parameters, Adam moments and step counters, EMA, density grid,
    bitfield, rng streams, counters).  Replayed steps see the same rays and samples; parameters agree up to the summation order of the
    floating-point atomics."""
    f = orc_flags(no_albedo=0, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=small_scene, flags=f, rays_per_batch=512)
    t.set_train_state(0, 512)
    for _ in range(20):
        t.train(want_stats=False)
    t.checkpoint_save()
    a = [t.train() for _ in range(4)]
    pa = t.get_params().copy(); st_a = t.get_train_state(); rng_a = t.get_rng()
    t.checkpoint_restore()
    b = [t.train() for _ in range(4)]
    pb = t.get_params(); st_b = t.get_train_state(); rng_b = t.get_rng()
    assert st_a == st_b and rng_a == rng_b
    for x, y in zip(a, b):
        assert x.n_samples == y.n_samples and x.n_rays_kept == y.n_rays_kept and x.training_step == y.training_step
        assert abs(x.n_samples_compacted - y.n_samples_compacted) <= 2
        assert abs(x.loss - y.loss) <= 1e-3 * abs(x.loss) + 1e-7
    assert rel_err(pb, pa) < 1e-4


def test_sdf_on_grid_matches_oracle(pkg):
    """rnb_sdf_on_grid (Testbed::get_density_on_grid, testbed_nerf.cu:4218-4269): lattice positions generated inside the tcgen05 probe
    kernel; against the oracle at the same lattice points and against rnb_eval_sdf on explicit positions."""
    import torch
    o, t = make_pair(pkg, MID, seed_params=7)
    res = (40, 36, 33); mn = np.array([0.1, 0.05, 0.2], np.float32); mx = np.array([0.9, 0.95, 0.85], np.float32)
    out = torch.empty(res[0] * res[1] * res[2], device="cuda")
    t.sdf_on_grid_device(res, mn, mx, out.data_ptr(), use_ema=False)
    torch.cuda.synchronize()
    iz, iy, ix = np.meshgrid(np.arange(res[2]), np.arange(res[1]), np.arange(res[0]), indexing="ij")       # x fastest
    idx = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], 1).astype(np.float32)
    inv = (np.float32(1.0) / np.array(res, np.float32)).astype(np.float32)
    pos = ((idx * inv) * (mx - mn) + mn).astype(np.float32)
    s_ref, _ = o.eval_sdf(pos, o.valid_level(0))
    got = out.cpu().numpy()
    assert rel_err(got, s_ref) < TOL
    x = torch.from_numpy(pos).cuda(); s2 = torch.empty(pos.shape[0], device="cuda")
    t.eval_sdf_device(x.data_ptr(), pos.shape[0], s2.data_ptr(), None, None, use_ema=False)
    torch.cuda.synchronize()
    assert half_close(got, s2.cpu().numpy(), ulps=1.0).mean() > 0.999
