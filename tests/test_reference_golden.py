"""Oracle (and, on a GPU, the CUDA path) against golden vectors produced by the UNMODIFIED reference.

tests/golden/ref_small*.npz were written by tests/ref_pin.py on a B200 box from oracle/_ref/bin/ref_harness — the reference's
own Testbed::train (src/testbed_nerf.cu, tiny-cuda-nn) compiled by oracle/Makefile.ref — on the 8-view 128x128 synthetic
scene, BASELINE configs[0] network (L=8, T=2^14, 32-wide MLPs), 256 rays/step, light index pinned to ray_idx % 3.

What is compared, and how tightly:
  * initial parameters (Trainer::initialize_params RNG order)                     : bit-exact
  * sample counts before / after compaction (marching, occupancy bits, T cut)     : exact integers
  * per-ray loss / Eikonal / mask terms (forward network, compositing, losses)    : 1e-4 relative (sorted: the reference's
                                                                                    ray slots come from atomicAdd order)
  * network outputs at probe points (sdf, normal, albedo logits)                  : 1e-3 relative
  * Adam + EMA fed with the reference's own gradients                             : 1e-6 absolute
  * gradient buffer                                                               : direction (cosine) and a LOOSE norm bound.
    The reference accumulates weight gradients in binary16 inside CUTLASS split-k GEMMs and grid gradients with binary16
    atomics (tcnn cutlass_matmul.h:83-84, grid.h:412-417); measured on the box (tests/golden/ref_pin_summary_*.json) its
    weight gradients are 3-7 % short of the fp32 sum and its coarse-level grid gradients carry ~10 % rounding noise, so
    1e-3 is not attainable against the reference's own arithmetic — the oracle's gradients are instead checked against
    finite differences in tests/test_oracle_gradcheck.py.
"""
import os
import numpy as np
import pytest

from oracle_binding import Oracle, default_flags
from common import SMALL, product_config, copy_flags, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip("golden file %s not present" % name)
    return np.load(p)


def _views(scene_mod):
    return scene_mod.make_scene(8, 128, 128, with_albedo=True)


def _sdf_init(g=None):
    """Geometric initialisation of the SDF MLP (the reference reads utils/mlp_weights*.txt, nerf_network.h:585-623): the 1536
    values the small config uses are the SDF-MLP slice of the reference's own initial parameters stored in the golden file."""
    g = g if g is not None else _load("ref_small.npz")
    return np.asarray(g["step0__params_in_mlp"][:32 * 32 + 16 * 32], np.float32)


def h2f(a):
    return np.asarray(a).view(np.float16).astype(np.float32)


def cos(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def sorted_rel(a, b):
    a = np.sort(np.asarray(a, np.float64)); b = np.sort(np.asarray(b, np.float64))
    assert a.size == b.size
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


class _Impl:
    """Uniform driver for the oracle and the CUDA Testbed."""

    def __init__(self, kind, pkg, views, albedo):
        self.kind = kind
        flags = default_flags(no_albedo=0 if albedo else 1, light_mode=-2)
        if kind == "oracle":
            self.o = Oracle(threads=min(8, os.cpu_count() or 1), **SMALL)
            self.o.set_flags(flags); self.o.set_views(views)
        else:
            self.lay = Oracle(threads=1, **SMALL)
            self.t = pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=256, pin_rays_per_batch=1))
            self.t.set_flags(copy_flags(pkg, flags)); self.t.load_training_data(views)

    @property
    def layout(self):
        return self.o if self.kind == "oracle" else self.lay

    def init_params(self, sdf_init):
        if self.kind == "oracle":
            self.o.init_params(1337, sdf_init); return self.o.get_params()[0]
        self.t.init_params(sdf_init); return self.t.get_params()

    def step(self, params, st, rays, bitfield=None):
        """One Testbed::train from the reference's in-state; returns dict(samples, compacted, loss, ek, mask, grads)."""
        if self.kind == "oracle":
            o = self.o
            o.set_params(params)
            if bitfield is not None:
                o.set_bitfield(bitfield)
            o.set_density_grid(np.zeros(128 ** 3, np.float32), int(st[8])) if bitfield is None else None
            o.set_train_state(training_step=int(st[4]), rays_per_batch=rays, n_rays_total=int(st[6]), measured_before=0, pin_rays=1)
            o.set_rng(int(st[0]), int(st[1]), int(st[2]), int(st[3]))
            s = o.train_step()
            _, lo, ek, ml = o.last_losses()
            return dict(samples=int(s.n_samples), compacted=int(s.n_compacted), loss=lo, ek=ek, mask=ml, grads=o.get_grads())
        t = self.t
        t.set_params(params)
        if bitfield is not None:
            t.set_bitfield(bitfield)
        else:
            t.import_density_grid(np.zeros(128 ** 3, np.float32), int(st[8]))
        t.set_train_state(int(st[4]), rays, int(st[6]), 0)
        t.set_rng([int(st[0]), int(st[1]), int(st[2]), int(st[3])])
        ts = int(st[4]); skip = min(max(ts // 16, 1), 16)
        if ts % skip == 0:
            t.training_prep_nerf()
        t.train_step_begin()
        g = t.get_grads(); _, l3 = t.ray_losses()
        s = t.train_step_end()
        return dict(samples=int(s.n_samples), compacted=int(s.n_samples_compacted), loss=l3[:, 0], ek=l3[:, 1], mask=l3[:, 2], grads=g)


def _check_all(kind, pkg, scene_mod, albedo):
    g = _load("ref_small_alb.npz" if albedo else "ref_small.npz")
    views = _views(scene_mod)
    impl = _Impl(kind, pkg, views, albedo)
    L = impl.layout
    nm = L.off_grid
    # ---- initialisation: same RNG order as Trainer::initialize_params / NerfNetwork::initialize_params ----
    p0 = impl.init_params(_sdf_init())
    assert np.array_equal(p0[:nm], g["step0__params_in_mlp"]), "MLP initialisation differs from the reference"
    idx = g["step0__params_in_grid_idx"].astype(np.int64)
    assert np.abs(p0[idx] - g["step0__params_in_grid_val"]).max() <= 1e-11
    gs = g["step0__params_in_grid_sum"]
    grid = p0[L.off_grid:L.off_var].astype(np.float64)
    assert abs(grid.sum() - gs[0]) <= 1e-9 * max(1.0, abs(gs[0])) + 1e-9 and abs((grid ** 2).sum() - gs[1]) <= 1e-9 * gs[1]
    # ---- step 0 from the initial state ----
    st = g["step0__state"]; cnt = g["step0__ref_counters"]
    r = impl.step(p0, st, int(cnt[0]))
    assert r["samples"] == int(cnt[1]) and r["compacted"] == int(cnt[2])
    assert sorted_rel(r["loss"], g["step0__ref_loss_sorted"]) < 1e-4
    assert sorted_rel(r["ek"], g["step0__ref_ek_sorted"]) < 1e-4
    assert sorted_rel(r["mask"], g["step0__ref_mask_sorted"]) < 1e-4
    ref_g = h2f(g["step0__ref_grads_fp16"])
    sdf = slice(L.off_sdf, L.off_rgb)
    assert cos(r["grads"][sdf], ref_g[sdf]) > 0.995 and rel_err(r["grads"][sdf], ref_g[sdf]) < 0.15
    assert np.count_nonzero(ref_g[L.off_grid:L.off_var]) == 0 and np.count_nonzero(r["grads"][L.off_grid:L.off_var]) == 0     # the reference trains no grid level at step 0
    if albedo:
        rgb = slice(L.off_rgb, L.off_grid)
        assert cos(r["grads"][rgb], ref_g[rgb]) > 0.95
    # ---- a later step (no occupancy refresh due): state = reference parameters + bitfield ----
    tag = [k.split("__")[0] for k in g.files if k.endswith("__params_in_fp16")][0]
    st = g[tag + "__state"]; cnt = g[tag + "__ref_counters"]
    assert int(st[4]) % min(max(int(st[4]) // 16, 1), 16) != 0
    pin = g[tag + "__params_in_fp16"].astype(np.float32)
    r = impl.step(pin, st, int(cnt[0]), bitfield=g[tag + "__bitfield"])
    assert r["samples"] == int(cnt[1])
    assert abs(r["compacted"] - int(cnt[2])) <= 2          # T < 1e-4 cut: a borderline sample may fall on either side
    tol = 1e-4 if r["compacted"] == int(cnt[2]) else 5e-3
    assert sorted_rel(r["loss"], g[tag + "__ref_loss_sorted"]) < tol
    assert sorted_rel(r["mask"], g[tag + "__ref_mask_sorted"]) < tol
    ref_g = h2f(g[tag + "__ref_grads_fp16"])
    assert cos(r["grads"][sdf], ref_g[sdf]) > 0.99
    gsl = slice(L.off_grid, L.off_var)
    assert np.count_nonzero(ref_g[gsl]) > 0
    assert cos(r["grads"][gsl], ref_g[gsl]) > 0.97 and 0.9 < np.linalg.norm(r["grads"][gsl]) / np.linalg.norm(ref_g[gsl]) < 1.1
    # same set of touched hash entries (index arithmetic): every entry the reference touched is touched here
    ref_nz = ref_g[gsl] != 0
    assert np.count_nonzero(ref_nz & (r["grads"][gsl] == 0)) <= 0.002 * np.count_nonzero(ref_nz)
    return impl, g


def test_oracle_matches_reference_golden(pkg, scene_mod):
    impl, g = _check_all("oracle", pkg, scene_mod, albedo=False)
    o = impl.o
    # ---- Adam + EMA on the reference's own gradients (adam.h:51-202, ema.h:116-152) ----
    oa = Oracle(threads=1, **SMALL)
    p0 = oa.get_params()[0] * 0
    oa.init_params(1337, _sdf_init()); p0 = oa.get_params()[0]
    oa.set_grads(h2f(g["step0__ref_grads_fp16"])); oa.optimizer_step()
    pa, _, ea = oa.get_params()
    assert np.abs(pa[:o.off_grid] - g["step0__ref_params_out_mlp"]).max() < 1e-6
    assert np.abs(pa[o.off_var:] - g["step0__ref_params_out_var"]).max() < 1e-6
    assert np.abs(ea[:o.off_grid] - h2f(g["step0__ref_ema_out_mlp_fp16"])).max() < 2e-3 * np.abs(ea[:o.off_grid]).max()
    # ---- network probe: NerfNetwork::inference_mixed_precision at 512 positions ----
    o.set_params(g["probe_params"].astype(np.float32))
    vl = o.valid_level(int(g["probe_state"][4]))
    out, _ = o.network_forward(g["probe_coords"], vl)
    ref = h2f(g["probe_out_fp16"]).reshape(-1, 16)
    assert rel_err(out[:, 3], ref[:, 3]) < 1e-3            # sdf
    assert rel_err(out[:, 4:7], ref[:, 4:7]) < 1e-3        # analytic normal
    assert np.array_equal(out[:, 7], ref[:, 7])            # variance slot


def test_oracle_matches_reference_golden_albedo(pkg, scene_mod):
    _check_all("oracle", pkg, scene_mod, albedo=True)


@pytest.mark.gpu
def test_cuda_matches_reference_golden(pkg, scene_mod):
    impl, g = _check_all("cuda", pkg, scene_mod, albedo=False)
    t = impl.t
    # Adam + EMA on the reference's gradients
    ta = pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=256, pin_rays_per_batch=1))
    ta.init_params(_sdf_init())
    ta.stage_optimizer(h2f(g["step0__ref_grads_fp16"]))
    pa = ta.get_params()
    assert np.abs(pa[:ta.off_grid] - g["step0__ref_params_out_mlp"]).max() < 1e-6
    # probe
    t.set_params(g["probe_params"].astype(np.float32))
    t.set_train_state(int(g["probe_state"][4]), 256, 0, 0)
    out, _ = t.stage_forward(g["probe_coords"])
    ref = h2f(g["probe_out_fp16"]).reshape(-1, 16)
    assert rel_err(out[:, 3], ref[:, 3]) < 1e-3 and rel_err(out[:, 4:7], ref[:, 4:7]) < 1e-3


@pytest.mark.gpu
def test_cuda_matches_reference_golden_albedo(pkg, scene_mod):
    _check_all("cuda", pkg, scene_mod, albedo=True)


# ---- default network (L=14, T=2^19, 64-wide) with ALL 14 hash levels live: reference outputs at 1024 probe points after 700 training steps of the
# reference build on B200 (tests/ref_pin.py --config full --steps 700 -> tests/golden/ref_full_probe.npz: MLP weights + only the hash entries the probes read)
FULL_PROBE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_full_probe.npz")


def _full_probe_params(o, g):
    p = np.zeros(o.n_params, np.float32)
    p[:o.off_grid] = g["mlp_fp16"].astype(np.float32)
    p[o.off_grid + g["grid_idx"].astype(np.int64)] = g["grid_val_fp16"].astype(np.float32)
    p[o.off_var:] = g["var"]
    return p


def _probe_errors(out, ref):
    return {"albedo_raw": rel_err(out[:, 0:3], ref[:, 0:3]), "sdf": rel_err(out[:, 3], ref[:, 3]), "normal": rel_err(out[:, 4:7], ref[:, 4:7]), "variance": rel_err(out[:, 7], ref[:, 7])}


@pytest.mark.skipif(not os.path.exists(FULL_PROBE), reason="tests/golden/ref_full_probe.npz not generated yet")
def test_oracle_matches_reference_full_network_all_levels_live():
    from oracle_binding import Oracle
    from common import FULL
    g = np.load(FULL_PROBE)
    vl = int(g["valid_level"][0])
    assert vl >= 13                                            # valid_level >= 13: all 14 levels enabled (grid.h:193-210, 1430-1437)
    o = Oracle(threads=4, **FULL)
    o.set_params(_full_probe_params(o, g))
    out, _ = o.network_forward(g["coords"], vl)
    ref = g["ref_out_fp16"].view(np.float16).astype(np.float32)
    e = _probe_errors(out, ref)
    # the reference accumulates every layer in binary16 (wmma, OUT_T = __half), the oracle in fp32 with binary16 rounding per layer: SDF / normal within
    # the north_star tolerance, the albedo logits (three 64-wide layers deep) at the binary16 floor
    assert e["sdf"] < 1e-3 and e["normal"] < 1e-3 and e["variance"] == 0.0, e
    assert e["albedo_raw"] < 2.5e-3, e


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(FULL_PROBE), reason="tests/golden/ref_full_probe.npz not generated yet")
def test_cuda_matches_reference_full_network_all_levels_live(pkg):
    from common import FULL, product_config
    g = np.load(FULL_PROBE)
    t = pkg.Testbed(product_config(pkg, FULL))
    class _O: pass
    o = _O(); o.n_params = t.n_params; o.off_grid = t.off_grid; o.off_var = t.off_var
    t.set_params(_full_probe_params(o, g))
    t.set_train_state(int(g["state"][4]), 256)
    out, _ = t.stage_forward(g["coords"])
    ref = g["ref_out_fp16"].view(np.float16).astype(np.float32)
    e = _probe_errors(out, ref)
    assert e["sdf"] < 1e-3 and e["normal"] < 1e-3 and e["variance"] == 0.0, e
    assert e["albedo_raw"] < 2.5e-3, e
