"""Size-independent properties of the CUDA path at BASELINE.json's full network size (L=14, T=2^19, 64-wide MLPs, 4096
rays/step), where the CPU oracle is too slow to be the checker.  Through the C ABI.

  * marching: every emitted sample lies in a cell whose occupancy bit is set (Morton index, bit-exact), samples of a ray
    advance strictly along the ray in whole multiples of dt = sqrt(3)/1024, sample slots are the exclusive prefix of the counts;
  * idempotence / determinism: the same state marched twice gives identical samples; the network forward is deterministic;
  * compaction: kept <= marched per ray, the compacted total is the sum of the kept counts, truncated at 2^18;
  * linearity of the backward in dL/dout (second-order terms included): g(2 d) = 2 g(d), g(d1 + d2) = g(d1) + g(d2);
  * sparse optimizer rule: a step with zero gradients leaves every hash-grid entry and its Adam state untouched, and moves
    the MLP weights only by weight decay;
  * a full training step lowers the loss on the batch it was taken on (sanity of the gradient sign at full size).
"""
import numpy as np
import pytest
from common import FULL, product_config

pytestmark = pytest.mark.gpu
DT = np.float32(1.73205080757) / np.float32(1024.0)


def _morton(ix, iy, iz):
    def part(v):
        v = v.astype(np.uint32)
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF); v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3); v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    return part(ix) | (part(iy) << np.uint32(1)) | (part(iz) << np.uint32(2))


@pytest.fixture(scope="module")
def trained(pkg, scene_mod):
    views = scene_mod.make_scene(12, 256, 256, with_albedo=True)
    t = pkg.Testbed(product_config(pkg, FULL, rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    t.init_params(); t.load_training_data(views)
    for _ in range(40):                       # past the first occupancy refreshes: a non-trivial bitfield
        t.train(want_stats=False)
    return t


def test_marching_invariants_full_size(trained):
    t = trained
    bf = t.get_bitfield()
    bits_all = np.unpackbits(bf, bitorder="little")            # 8 mips x 128^3 bits
    bits = bits_all[:128 ** 3]
    assert 0 < bits.sum() < 128 ** 3
    ts, R, nrt, _ = t.get_train_state()
    a = t.stage_generate(4096, nrt, 1 << 22)
    b = t.stage_generate(4096, nrt, 1 << 22)
    K = a["n_kept"]
    assert K > 1000 and a["n_samples"] == b["n_samples"] and np.array_equal(a["numsteps"], b["numsteps"])
    n = int(a["n_samples"])
    assert np.array_equal(a["coords"][:n].view(np.uint32), b["coords"][:n].view(np.uint32))          # idempotent, bit for bit
    ns = a["numsteps"]
    assert np.array_equal(ns[:, 1], np.concatenate([[0], np.cumsum(ns[:-1, 0])]).astype(np.uint32))   # slots = exclusive prefix of the counts
    assert ns[:, 0].max() <= 1024 and ns[:, 0].min() >= 1
    pos = a["coords"][:n, :3]
    assert pos.min() >= 0.0 and pos.max() <= 1.0
    # cascaded_grid_idx_at (testbed_nerf.cu:439-459) in the same fp32 operation order; mip_from_pos (:569-574): a coordinate exactly
    # on the cube face (|x - 0.5| == 0.5) selects mip 1
    c = pos - np.float32(0.5)
    mx = np.abs(c).max(axis=1)
    _, e = np.frexp(mx)
    mip = np.clip(e + 1, 0, 7).astype(np.int64)
    ms = np.ldexp(np.float32(1.0), -mip).astype(np.float32)[:, None]
    cell = np.clip(((c * ms + np.float32(0.5)) * np.float32(128.0)).astype(np.int32), 0, 127)
    idx = _morton(cell[:, 0], cell[:, 1], cell[:, 2]).astype(np.int64) + mip * (128 ** 3)
    assert (mip <= 1).all() and (mip == 0).mean() > 0.999
    bad = np.flatnonzero(bits_all[idx] == 0)
    assert bad.size == 0, ("samples emitted in unoccupied cells", bad.size, pos[bad[:5]], mip[bad[:5]])
    # along each ray: consecutive samples are k * dt apart, k >= 1 integer (skips are whole steps), direction = ray direction
    ray_of = np.repeat(np.arange(K), ns[:, 0])
    same = ray_of[1:] == ray_of[:-1]
    d = a["rays"][:, 3:6]; d = d / np.linalg.norm(d, axis=1, keepdims=True)
    step = ((pos[1:] - pos[:-1]) * d[ray_of[1:]]).sum(1)[same] / DT
    assert (step > 0.5).all()
    assert np.abs(step - np.rint(step)).max() < 2e-2


def test_forward_deterministic_and_compaction_counts(trained):
    t = trained
    s = t.train()
    marched, kept = t.ray_counts()
    # n_samples is the reference's numsteps counter: it also counts rays that did not fit the sample budget of the step (:1353-1357)
    assert (kept <= marched).all() and int(marched.sum()) <= s.n_samples
    assert int(kept.sum()) == s.n_samples_compacted
    assert s.n_samples_trained == min(s.n_samples_compacted, 1 << 18)
    nrt = t.get_train_state()[2]
    g = t.stage_generate(512, nrt, 1 << 20)
    n = int(g["n_samples"])
    o1, _ = t.stage_forward(g["coords"][:min(n, 20000)], want_normal=False)
    o2, _ = t.stage_forward(g["coords"][:min(n, 20000)], want_normal=False)
    assert np.array_equal(o1, o2)
    assert np.isfinite(o1).all()


def test_backward_is_linear_in_dout(trained):
    t = trained
    nrt = t.get_train_state()[2]
    g = t.stage_generate(256, nrt, 1 << 20)
    n = min(int(g["n_samples"]), 16384)
    coords = g["coords"][:n]
    rs = np.random.RandomState(5)
    # dL/dout values on a binary16-friendly grid (multiples of 2^-10 in [-1/8, 1/8]) so that scaling by two and adding stay exact in binary16
    def rnd():
        d = np.zeros((n, 16), np.float32)
        d[:, :11] = rs.randint(-128, 129, size=(n, 11)).astype(np.float32) / 1024.0
        d[:, 7] = 0.0
        return d
    d1, d2 = rnd(), rnd()
    g1 = t.stage_backward(coords, d1, n); g2 = t.stage_backward(coords, d2, n)
    g12 = t.stage_backward(coords, d1 + d2, n); g1x2 = t.stage_backward(coords, 2 * d1, n)
    def rel(a, b):
        return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
    assert np.linalg.norm(g1) > 0
    # fp16 rounding of intermediate gradients makes this approximate: 2e-3 norm-wise (one binary16 ulp is 1e-3)
    assert rel(g1x2, 2.0 * g1.astype(np.float64)) < 2e-3
    assert rel(g12, g1.astype(np.float64) + g2) < 2e-3
    # the touched hash entries depend on the positions only
    grid = slice(t.off_grid, t.off_var)
    assert np.array_equal(g1x2[grid] != 0, g1[grid] != 0)


def test_sparse_optimizer_rule_full_size(trained):
    t = trained
    p0 = t.get_params().copy()
    t.stage_optimizer(np.zeros(t.n_params, np.float32))
    p1 = t.get_params()
    grid = slice(t.off_grid, t.off_var)
    assert np.array_equal(p0[grid], p1[grid]), "hash entries with zero gradient must not move (adam.h:111-115)"
    mlp = slice(0, t.off_grid)
    assert np.abs(p1[mlp] - p0[mlp]).max() < 5e-3         # weight decay / momentum only
    t.set_params(p0)


def test_training_step_lowers_batch_loss(pkg, scene_mod):
    views = scene_mod.make_scene(12, 256, 256, with_albedo=True)
    t = pkg.Testbed(product_config(pkg, FULL, rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    t.init_params(); t.load_training_data(views)
    losses = [t.train().loss for _ in range(60)]
    assert np.isfinite(losses).all()
    assert np.mean(losses[-10:]) < 0.75 * np.mean(losses[:5]), losses[::6]      # measured: 0.55 x after 60 steps
