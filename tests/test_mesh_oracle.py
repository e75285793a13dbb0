"""CPU tests of the mesh oracle (oracle/orc_mesh.h: sequential restatement of src/marching_cubes.cu gen_vertices / gen_faces /
accumulate_1ring / save_mesh).  The reference has no golden vectors for this path, so the oracle is pinned by
  * structure: one vertex per sign-changing lattice edge, a closed 2-manifold (every edge in exactly two triangles, V-E+F = 2)
    for a sphere, consistently oriented, vertices on the iso-surface of the trilinear field;
  * the text writer against an independent formatter: Python's '%0.5f' (correctly rounded, as glibc's) on numpy float32
    arithmetic in the reference's operation order;
and, on the B200 box, against the reference build's own mesh (tests/ref_pin.py, tests/golden/ref_pin_summary_*.json)."""
import os
import numpy as np
import pytest
import oracle_binding as ob


def sphere_field(res, radius=0.31, center=(0.5, 0.47, 0.52), mn=(0, 0, 0), mx=(1, 1, 1), bumps=0.0):
    rx, ry, rz = res
    z, y, x = np.meshgrid(np.arange(rz), np.arange(ry), np.arange(rx), indexing="ij")
    # lattice point -> position with the spacing the extraction assumes: lattice * (max - min) / res + min
    px = x / rx * (mx[0] - mn[0]) + mn[0]; py = y / ry * (mx[1] - mn[1]) + mn[1]; pz = z / rz * (mx[2] - mn[2]) + mn[2]
    d = np.sqrt((px - center[0]) ** 2 + (py - center[1]) ** 2 + (pz - center[2]) ** 2) - radius
    if bumps:
        d = d + bumps * np.sin(19 * px) * np.cos(23 * py) * np.sin(17 * pz)
    return d.astype(np.float32)


def edge_counts(F):
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    und = np.sort(e, axis=1)
    _, cnt = np.unique(und, axis=0, return_counts=True)
    return e, cnt


def test_sphere_is_a_closed_oriented_manifold():
    res = (32, 48, 40)
    d = sphere_field(res)
    V, N, I, nv = ob.marching_cubes(d, thresh=0.0)
    F = I.reshape(-1, 3).astype(np.int64)
    inside = d > 0
    crossings = (inside[:, :, 1:] != inside[:, :, :-1]).sum() + (inside[:, 1:, :] != inside[:, :-1, :]).sum() + (inside[1:] != inside[:-1]).sum()
    assert nv == crossings and nv > 1000
    assert V.shape[0] == (nv + 127) // 128 * 128 and np.all(V[nv:] == 0) and np.all(N[nv:] == 0)
    assert F.min() >= 0 and F.max() == nv - 1 and np.unique(F).size == nv               # every vertex is used
    e, cnt = edge_counts(F)
    assert np.all(cnt == 2)                                                              # closed 2-manifold
    assert nv - cnt.size + F.shape[0] == 2                                               # genus 0
    # consistent orientation: every directed edge appears once in each direction
    key = e[:, 0] * (nv + 1) + e[:, 1]; rkey = e[:, 1] * (nv + 1) + e[:, 0]
    assert np.array_equal(np.sort(key), np.sort(rkey)) and np.unique(key).size == key.size
    # vertices sit on the sphere up to the interpolation error of one cell; summed face normals are radial (all the same way)
    c = np.array([0.5, 0.47, 0.52], np.float32)
    r = np.linalg.norm(V[:nv] - c, axis=1)
    assert np.abs(r - 0.31).max() < 2e-3
    radial = ((V[:nv] - c) * N[:nv]).sum(1) / (np.linalg.norm(N[:nv], axis=1) * r)
    assert np.all(np.abs(radial) > 0.9) and (np.all(radial > 0) or np.all(radial < 0))


def test_vertex_positions_and_order_follow_the_lattice():
    res = (16, 12, 8); mn = (-1.0, 0.25, 2.0); mx = (3.0, 1.0, 2.5)
    rs = np.random.RandomState(3)
    d = rs.uniform(-1, 1, (res[2], res[1], res[0])).astype(np.float32)               # every cube case shows up
    th = np.float32(0.1)
    V, N, I, nv = ob.marching_cubes(d, mn, mx, thresh=float(th))
    # independent numpy walk in the same order: point (x fastest), then its +x, +y, +z edge
    s = ((np.array(mx, np.float32) - np.array(mn, np.float32)) / np.array(res, np.float32)).astype(np.float32)
    exp = []
    for z in range(res[2]):
        for y in range(res[1]):
            for x in range(res[0]):
                f0 = d[z, y, x]
                for a, (dx, dy, dz) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
                    if x + dx >= res[0] or y + dy >= res[1] or z + dz >= res[2]:
                        continue
                    f1 = d[z + dz, y + dy, x + dx]
                    if (f0 > th) == (f1 > th):
                        continue
                    dt = np.float32((th - f0) / np.float32(f1 - f0))
                    l = np.array([x, y, z], np.float32); l[a] = np.float32(l[a] + dt)
                    exp.append(np.array([np.float64(l[k]) * np.float64(s[k]) + np.float64(mn[k]) for k in range(3)]).astype(np.float32))   # fma: one rounding
    exp = np.array(exp, np.float32)
    assert nv == exp.shape[0]
    assert np.array_equal(V[:nv].view(np.uint32), exp.view(np.uint32))
    F = I.reshape(-1, 3)
    assert F.max() < nv and I.size % 3 == 0 and I.size > 0


def _fmt_expected(verts, normals, colors, idx, scale, off, s, t, invert):
    f32 = np.float32
    lines = []
    for v, c in zip(verts, colors):
        p = [f32(f32(s) * f32(f32(v[d] - f32(off[d])) / f32(scale))) + f32(t[d]) for d in range(3)]
        cc = [c[d] if np.isnan(c[d]) else min(max(c[d], f32(0)), f32(1)) for d in range(3)]
        lines.append("v %0.5f %0.5f %0.5f %0.3f %0.3f %0.3f" % (*[float(x) for x in p], *[float(x) for x in cc]))
    for n in normals:
        m = [f32(f32(s) * n[d]) for d in range(3)]
        z = f32(f32(m[0] * m[0]) + f32(f32(m[1] * m[1]) + f32(m[2] * m[2])))
        if z > 0:
            r = np.sqrt(z, dtype=f32); m = [f32(x / r) for x in m]
        lines.append("vn %0.5f %0.5f %0.5f" % tuple(float(x) for x in m))
    for a, b, c in idx.reshape(-1, 3):
        if not invert:
            a, c = c, a
        lines.append("f %d//%d %d//%d %d//%d" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("invert", [False, True])
def test_obj_writer_matches_independent_formatter(tmp_path, invert):
    rs = np.random.RandomState(11)
    n = 600
    with np.errstate(all="ignore"):
        V = rs.uniform(-2, 2, (n, 3)).astype(np.float32)
        # rounding ties and near-ties at the 5th decimal, signed zeros, tiny and large magnitudes
        V[:50, 0] = (np.arange(50) * 2 + 1) * np.float32(0.5e-5); V[50:60, 1] = np.float32(-1e-9); V[60:70, 2] = np.float32(123456.789)
        V[70:80, 0] = np.float32(-0.0); V[80:90, 1] = np.float32(1e-42); V[90:100, 2] = np.float32(0.999995)
        N = rs.normal(0, 1e-4, (n, 3)).astype(np.float32); N[:20] = 0
        Cc = rs.uniform(-0.2, 1.2, (n, 3)).astype(np.float32); Cc[:30, 0] = (np.arange(30) * 2 + 1) * np.float32(0.5e-3)
        I = rs.randint(0, n, 3 * 400).astype(np.uint32)
        scale, off, s, t = 0.5, (0.5, 0.5, 0.5), 1.7, (0.25, -3.0, 10.0)
        path = tmp_path / "m.obj"
        ob.save_mesh(path, V, N, Cc, I, scale, off, s, t, invert)
        got = open(path).read()
        exp = _fmt_expected(V, N, Cc, I, scale, off, s, t, invert)
    assert got == exp


def test_ply_writer_layout(tmp_path):
    rs = np.random.RandomState(5)
    V = rs.uniform(0, 1, (8, 3)).astype(np.float32); N = rs.normal(0, 1, (8, 3)).astype(np.float32); Cc = rs.uniform(-0.1, 1.1, (8, 3)).astype(np.float32)
    I = np.array([0, 1, 2, 2, 3, 4], np.uint32)
    path = tmp_path / "m.ply"
    ob.save_mesh(path, V, N, Cc, I, 1.0, (0, 0, 0), 1.0, (0, 0, 0), False)
    txt = open(path).read().split("\n")
    assert txt[0] == "ply" and "element vertex 8" in txt and "element face 2" in txt
    body = txt[txt.index("end_header") + 1:]
    first = body[0].split()
    assert len(first) == 9 and first[:3] == ["%0.5f" % float(x) for x in V[0]]
    nn = N[0] / np.sqrt(N[0, 0] * N[0, 0] + (N[0, 1] * N[0, 1] + N[0, 2] * N[0, 2]))
    assert first[3:6] == ["%0.3f" % float(x) for x in nn]
    assert first[6:] == [str(int(np.uint8(min(max(float(c) * 255.0, 0.0), 255.0)))) for c in (Cc[0] * np.float32(1)).astype(np.float32)]
    assert body[8] == "3 2 1 0" and body[9] == "3 4 3 2"
