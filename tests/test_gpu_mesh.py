"""Mesh extraction and export on the GPU through the C ABI, against the CPU oracle (oracle/orc_mesh.h).

Integer work is bit-exact: vertex numbering, triangle indices, the bytes of the OBJ / PLY text.  Vertex positions and the
area-weighted normals are bit-exact as well because the kernels use the lattice order and the summation order of the
sequential restatement.  Vertex colours go through the binary16 network: 2e-3 absolute after the logistic."""
import numpy as np
import pytest
import oracle_binding as ob
from common import MID, FULL, make_pair, product_config
from test_mesh_oracle import sphere_field, edge_counts

pytestmark = pytest.mark.gpu


def _extract(t, d, mn, mx, th, with_colors=False):
    import torch
    dev = torch.from_numpy(np.ascontiguousarray(d, np.float32)).cuda()
    info = t.marching_cubes_from_density(dev.data_ptr(), (d.shape[2], d.shape[1], d.shape[0]), mn, mx, th, with_colors=with_colors, use_ema=False)
    torch.cuda.synchronize()
    return info, t.mesh_download()


@pytest.fixture(scope="module")
def tb(pkg):
    t = pkg.Testbed(product_config(pkg, MID))
    t.init_params()
    return t


@pytest.mark.parametrize("case", ["sphere", "bumpy", "noise", "noise48", "empty", "thin"])
def test_extraction_matches_oracle_bit_for_bit(tb, case):
    mn, mx, th = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), 0.0
    if case == "sphere":
        d = sphere_field((64, 48, 40))
    elif case == "bumpy":
        mn, mx, th = (-0.5, 0.1, 0.0), (1.5, 0.9, 2.0), 0.013
        d = sphere_field((48, 64, 36), radius=0.33, center=(0.5, 0.5, 1.0), mn=mn, mx=mx, bumps=0.05)
    elif case == "noise":                      # all 256 cube cases, surface through the lattice boundary
        th = 0.05
        d = np.random.RandomState(1).uniform(-1, 1, (24, 20, 32)).astype(np.float32)
    elif case == "noise48":                    # rows of 48: the 16-point groups straddle the 32-bit words of the sign-bit array
        th = -0.1
        d = np.random.RandomState(2).uniform(-1, 1, (7, 9, 48)).astype(np.float32)
    elif case == "empty":
        d = np.full((8, 8, 16), 1.0, np.float32)
    else:                                      # smallest lattice the API takes, one crossing plane
        d = np.zeros((2, 3, 16), np.float32) - 1.0; d[:, :, 9:] = 1.0
    info, m = _extract(tb, d, mn, mx, th)
    V, N, I, nv = ob.marching_cubes(d, mn, mx, th)
    assert info["n_verts"] == nv and info["n_verts_padded"] == V.shape[0] and info["n_indices"] == I.size
    assert np.array_equal(m["F"].ravel(), I)
    assert np.array_equal(m["V"].view(np.uint32), V.view(np.uint32))
    assert np.array_equal(m["N"].view(np.uint32), N.view(np.uint32))
    assert np.all(m["C"] == 0)
    if case == "empty":
        assert nv == 0 and I.size == 0


def test_extraction_is_deterministic_and_closed_at_256(tb):
    d = sphere_field((256, 256, 256), radius=0.4, bumps=0.02)
    _, a = _extract(tb, d, (0, 0, 0), (1, 1, 1), 0.0)
    _, b = _extract(tb, d, (0, 0, 0), (1, 1, 1), 0.0)
    for k in ("V", "N", "F"):
        assert np.array_equal(a[k], b[k])
    nv = a["n_verts"]; F = a["F"].astype(np.int64)
    assert nv > 150000 and F.max() == nv - 1
    e, cnt = edge_counts(F)
    assert np.all(cnt == 2) and nv - cnt.size + F.shape[0] == 2


def _tricky_mesh(n=5000, seed=11):
    rs = np.random.RandomState(seed)
    with np.errstate(all="ignore"):
        V = rs.uniform(-2, 2, (n, 3)).astype(np.float32)
        V[:50, 0] = (np.arange(50) * 2 + 1) * np.float32(0.5e-5); V[50:60, 1] = np.float32(-1e-9); V[60:70, 2] = np.float32(123456.789)
        V[70:80, 0] = np.float32(-0.0); V[80:90, 1] = np.float32(1e-42); V[90:100, 2] = np.float32(0.999995)
        V[100:110, 0] = np.float32(3.0e38); V[110:120, 1] = np.float32(-7.5e20); V[120:125, 2] = np.float32(np.inf); V[125:130, 0] = np.float32(np.nan)
        V[130:140, 1] = np.float32(16777216.0); V[140:150, 2] = np.float32(8388607.5)
        N = rs.normal(0, 1e-4, (n, 3)).astype(np.float32); N[:20] = 0
        Cc = rs.uniform(-0.2, 1.2, (n, 3)).astype(np.float32); Cc[:30, 0] = (np.arange(30) * 2 + 1) * np.float32(0.5e-3); Cc[30:35, 1] = np.float32(np.nan)
        I = rs.randint(0, n, 3 * (n + 77)).astype(np.uint32)
    return V, N, Cc, I


@pytest.mark.parametrize("ext,invert", [("obj", False), ("obj", True), ("ply", False), ("ply", True)])
def test_mesh_text_is_byte_identical_to_fprintf(pkg, tmp_path, ext, invert):
    import torch
    V, N, Cc, I = _tricky_mesh()
    if ext == "ply":
        Cc = np.nan_to_num(Cc, nan=0.5)            # (unsigned char) of a NaN is undefined behaviour on the host side
    scale, off, s, t = 0.5, (0.5, 0.5, 0.5), 1.7, (0.25, -3.0, 10.0)
    ref = tmp_path / ("ref." + ext); got = tmp_path / ("got." + ext)
    ob.save_mesh(ref, V, N, Cc, I, scale, off, s, t, invert)
    dv, dn, dc, di = (torch.from_numpy(x).cuda() for x in (V, N, Cc, I.view(np.int32)))
    nb = pkg.save_mesh_device(got, dv.data_ptr(), dn.data_ptr(), dc.data_ptr(), di.data_ptr(), V.shape[0], I.size, scale, off, s, t, invert)
    a = open(ref, "rb").read(); b = open(got, "rb").read()
    assert nb == len(b)
    if a != b:
        la, lb = a.split(b"\n"), b.split(b"\n")
        bad = [(i, x, y) for i, (x, y) in enumerate(zip(la, lb)) if x != y][:5]
        raise AssertionError(("text differs", len(a), len(b), bad))


def test_marching_cubes_end_to_end_small(pkg, tmp_path):
    """Testbed::marching_cubes on a network: our SDF sweep -> extraction -> colours; colours against the oracle's network at the
    vertices (direction = normalised (p - 0.5), as generate_nerf_network_inputs_from_positions), file written and re-read."""
    import torch
    o, t = make_pair(pkg, MID, seed_params=None)
    info = t.marching_cubes(40, use_ema=False)              # rounded up to 48
    assert info["res"] == (48, 48, 48) and info["n_verts"] > 0
    m = t.mesh_download()
    # same mesh from the explicit two-step path
    sdf = torch.empty(48 ** 3, device="cuda")
    t.sdf_on_grid_device((48, 48, 48), (0, 0, 0), (1, 1, 1), sdf.data_ptr(), use_ema=False)
    torch.cuda.synchronize()
    d = sdf.cpu().numpy().reshape(48, 48, 48)
    V, N, I, nv = ob.marching_cubes(d, thresh=0.0)
    assert nv == info["n_verts"] and np.array_equal(m["V"], V) and np.array_equal(m["F"].ravel(), I) and np.array_equal(m["N"], N)
    # colours
    P = m["V"]; dirs = P - np.float32(0.5)
    z = (dirs * dirs).sum(1, keepdims=True); dirs = np.where(z > 0, dirs / np.sqrt(np.maximum(z, 1e-30)), dirs)
    coords = np.concatenate([P, np.zeros((P.shape[0], 1), np.float32), (dirs + 1) * 0.5], 1).astype(np.float32)
    out, _ = o.network_forward(coords, o.valid_level(0))
    exp = 1.0 / (1.0 + np.exp(-out[:, :3].astype(np.float64)))
    assert np.abs(m["C"] - exp).max() < 2e-3
    # file: vertex / normal / face line counts and a parse of the first vertex
    path = tmp_path / "mesh.obj"
    nb = t.save_mesh(path, nerf_scale=0.5, nerf_offset=(0.5, 0.5, 0.5), n2w_s=2.0, n2w_t=(1.0, 2.0, 3.0), invert_normals=True)
    txt = open(path).read()
    assert len(txt) == nb
    lines = txt.split("\n")
    nvp = info["n_verts_padded"]
    assert sum(l.startswith("v ") for l in lines) == nvp and sum(l.startswith("vn ") for l in lines) == nvp and sum(l.startswith("f ") for l in lines) == info["n_indices"] // 3
    ref = tmp_path / "ref.obj"
    ob.save_mesh(ref, m["V"], m["N"], m["C"], m["F"].ravel(), 0.5, (0.5, 0.5, 0.5), 2.0, (1.0, 2.0, 3.0), True)
    assert open(ref).read() == txt


def test_marching_cubes_full_size_network_256(pkg, scene_mod):
    """Default network (L=14, T=2^19) after a short training run: the 256^3 mesh of the SDF is a closed surface inside the cube."""
    views = scene_mod.make_scene(12, 256, 256, with_albedo=True)
    t = pkg.Testbed(product_config(pkg, FULL, rays_per_batch=4096, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    t.init_params(); t.load_training_data(views)
    for _ in range(60):
        t.train(want_stats=False)
    info = t.marching_cubes(256, use_ema=True)
    m = t.mesh_download()
    nv = info["n_verts"]; F = m["F"].astype(np.int64)
    assert nv > 10000 and F.max() == nv - 1
    assert np.isfinite(m["V"]).all() and m["V"][:nv].min() >= 0.0 and m["V"][:nv].max() <= 1.0
    e, cnt = edge_counts(F)
    assert np.all(cnt <= 2) and (cnt == 2).mean() > 0.99          # open only where the surface leaves the lattice
    assert np.isfinite(m["C"]).all() and m["C"].min() >= 0.0 and m["C"].max() <= 1.0
