"""CPU check of the ray/mesh traversal the CUDA kernel runs (rnb_raymesh.cuh compiled for the host) against brute force over all
triangles (oracle/orc_albedo.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc_albedo  # noqa: E402
from albedo_scene import icosphere  # noqa: E402


@pytest.fixture(scope="module")
def host_lib():
    src = os.path.join(ROOT, "tests", "cuda", "raymesh_host.cpp"); so = os.path.join(ROOT, "tests", "cuda", "libraymesh_host.so")
    deps = [src] + [os.path.join(ROOT, "rnb-neus2_b200", "csrc", f) for f in ("rnb_raymesh.cuh", "rnb_raymesh_build.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


def _trace(L, verts, tris, org, dirs, t_max=None, grid_res=0):
    n = org.shape[0]
    t = np.zeros(n); tri = np.zeros(n, np.uint32); res = (C.c_uint32 * 3)()
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty)) if a is not None else None
    org = np.ascontiguousarray(org, np.float64); dirs = np.ascontiguousarray(dirs, np.float64)
    tm = np.ascontiguousarray(t_max, np.float64) if t_max is not None else None
    L.raymesh_host_trace(p(verts, C.c_float), C.c_uint32(len(verts)), p(tris, C.c_uint32), C.c_uint32(len(tris)), C.c_uint32(grid_res),
                         p(org, C.c_double), p(dirs, C.c_double), p(tm, C.c_double), C.c_uint32(n), p(t, C.c_double), p(tri, C.c_uint32), res)
    return t, tri, tuple(res)


def two_spheres():
    v1, f1 = icosphere(3, 1.0); v2, f2 = icosphere(2, 0.35, (1.6, 0.4, 0.2))
    return np.ascontiguousarray(np.concatenate([v1, v2])), np.ascontiguousarray(np.concatenate([f1, f2 + len(v1)]).astype(np.uint32))


def random_rays(n, seed):
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3)); o = o / np.linalg.norm(o, axis=1)[:, None] * rng.uniform(2.0, 4.0, size=(n, 1))
    target = rng.uniform(-1.3, 1.3, size=(n, 3))
    d = target - o; d /= np.linalg.norm(d, axis=1)[:, None]
    o[: n // 8] = rng.uniform(-0.5, 0.5, size=(n // 8, 3))          # some origins inside the big sphere
    return o, d


@pytest.mark.parametrize("grid_res", [0, 4, 37, 128])
def test_first_hit_matches_bruteforce(host_lib, grid_res):
    verts, tris = two_spheres()
    o, d = random_rays(1500, 1)
    t, tri, res = _trace(host_lib, verts, tris, o, d, None, grid_res)
    t_ref, tri_ref = orc_albedo.first_hit(verts, tris, o, d)
    assert np.array_equal(np.isfinite(t), np.isfinite(t_ref))
    m = np.isfinite(t_ref)
    assert m.sum() > 700
    assert np.max(np.abs(t[m] - t_ref[m])) < 1e-9
    same = tri[m] == tri_ref[m]
    assert same.mean() > 0.995          # a ray through a shared edge may report either neighbour
    assert np.all(tri[~m] == orc_albedo.NO_TRI)


def test_axis_aligned_and_degenerate_rays(host_lib):
    verts, tris = two_spheres()
    o = np.array([[3.0, 0.0, 0.0], [0.0, -3.0, 0.0], [0.0, 0.0, 3.0], [3.0, 3.0, 3.0], [0.0, 0.0, 0.0], [5.0, 0.01, 0.02]])
    d = np.array([[-1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [-1.0, 0.0, 0.0]])
    t, tri, _ = _trace(host_lib, verts, tris, o, d)
    t_ref, _ = orc_albedo.first_hit(verts, tris, o, d)
    assert np.array_equal(np.isfinite(t), np.isfinite(t_ref))
    m = np.isfinite(t_ref)
    assert np.allclose(t[m], t_ref[m], atol=1e-9)
    assert not np.isfinite(t[3])          # pointing away from everything


def test_any_hit_matches_bruteforce(host_lib):
    verts, tris = two_spheres()
    o, d = random_rays(1500, 2)
    rng = np.random.default_rng(3)
    t_max = rng.uniform(0.5, 5.0, size=o.shape[0])
    _, tri, _ = _trace(host_lib, verts, tris, o, d, t_max)
    ref = orc_albedo.any_hit(verts, tris, o, d, t_max)
    got = tri != orc_albedo.NO_TRI
    assert np.array_equal(got, ref)
    assert 100 < ref.sum() < 1400


def test_unnormalised_directions_scale_the_parameter(host_lib):
    verts, tris = two_spheres()
    o, d = random_rays(300, 4)
    t1, tri1, _ = _trace(host_lib, verts, tris, o, d)
    t2, tri2, _ = _trace(host_lib, verts, tris, o, d * 2.5)
    m = np.isfinite(t1)
    assert np.array_equal(m, np.isfinite(t2))
    assert np.allclose(t1[m], t2[m] * 2.5, rtol=1e-12)


def test_c_abi_argument_checks_without_a_device(pkg):
    """rnb_raymesh_create validates its arguments before it touches the device, and has no CPU path behind it"""
    import torch
    L = pkg.lib()
    verts, tris = icosphere(1)
    h = C.c_void_p()
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    INVALID, CUDA = 1, None
    bad_idx = tris.copy(); bad_idx[3, 1] = len(verts)
    nan_v = verts.copy(); nan_v[5, 2] = np.nan
    cases = [(None, 0, None, 0), (verts, len(verts), None, 0), (verts, len(verts), tris, 0), (verts, 0, tris, len(tris)), (verts, len(verts), bad_idx, len(tris)), (nan_v, len(verts), tris, len(tris))]
    for v, nv, t, nt in cases:
        rc = L.rnb_raymesh_create(p(v, C.c_float) if v is not None else None, C.c_uint32(nv), p(t, C.c_uint32) if t is not None else None, C.c_uint32(nt), C.c_uint32(0), C.byref(h))
        assert rc != 0 and len(L.rnb_last_error()) > 0
    assert b"not finite" in L.rnb_last_error()
    if not torch.cuda.is_available():
        rc = L.rnb_raymesh_create(p(verts, C.c_float), C.c_uint32(len(verts)), p(tris, C.c_uint32), C.c_uint32(len(tris)), C.c_uint32(0), C.byref(h))
        assert rc != 0 and b"no CPU path" in L.rnb_last_error()
    assert L.rnb_raymesh_intersect(None, None, None, None, C.c_uint32(4), None, None, None) != 0
    assert L.rnb_raymesh_info(None, None, None) != 0
    assert L.rnb_raymesh_destroy(None) == 0


def test_large_triangles_coarsen_the_grid(host_lib):
    """60 triangles that span the whole box next to a finely tessellated object: at 320^3 cells each of them would be referenced from
    millions of cells (2 G references); the grid is coarsened until the list is bounded, and the answers stay those of brute force"""
    v1, f1 = icosphere(4, 0.05)                                              # 5120 small triangles
    rng = np.random.default_rng(9)
    big = rng.uniform(-40, 40, size=(180, 3)).astype(np.float32)
    verts = np.ascontiguousarray(np.concatenate([v1, big])); n = len(v1)
    tris = np.ascontiguousarray(np.concatenate([f1, n + np.arange(180).reshape(60, 3)]).astype(np.uint32))
    o = rng.uniform(-30, 30, size=(400, 3))
    tgt = rng.uniform(-5, 5, size=(400, 3)); tgt[:100] = rng.uniform(-0.05, 0.05, size=(100, 3))
    d = tgt - o; d /= np.linalg.norm(d, axis=1)[:, None]
    t, tri, res = _trace(host_lib, verts, tris, o, d, None, 320)
    assert max(res) <= 80, res                                               # 320 -> 160 -> 80: the bounding boxes of the 60 triangles then cover < 4 M cells
    t_ref, tri_ref = orc_albedo.first_hit(verts, tris, o, d)
    assert np.array_equal(np.isfinite(t), np.isfinite(t_ref)) and np.isfinite(t_ref).sum() > 100
    m = np.isfinite(t_ref)
    assert np.max(np.abs(t[m] - t_ref[m])) < 1e-8
    assert (tri[m] == tri_ref[m]).mean() > 0.99
