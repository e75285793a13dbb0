"""The mesh oracle against the REFERENCE's own mesh (CPU, no reference needed at run time).

tests/golden/ref_mesh_lattice32.npz was produced on the B200 box by tests/ref_pin_mesh.py (--config full --mesh 30 --golden
lattice32): the unmodified reference build (oracle/_ref/bin/ref_harness) trained the default network for 120 steps, ran get_density_on_grid and
compute_and_save_marching_cubes_mesh as src/main.cu:460 does, and the harness dumped the SDF lattice, m_mesh.* and the files.
The reference numbers vertices / triangles in atomicAdd arrival order, so meshes are compared as sets:
  * vertex positions bit-exact, triangles (ordered position triples up to rotation) bit-exact,
  * area-weighted normals to 1e-5 of their length (the reference sums them with float atomics),
  * the text writer on the reference's own arrays: byte-identical OBJ and PLY."""
import os
import numpy as np
import pytest
import oracle_binding as ob
from ref_pin_mesh import canonical_triangles, sort_rows, vertex_rows

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mesh_lattice32.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_marching_cubes_matches_reference_mesh_as_sets(gold):
    D = gold["density"]; mn, mx = (tuple(float(x) for x in r) for r in gold["aabb"])
    rV, rN, rI = gold["verts"], gold["normals"], gold["indices"]
    V, N, I, nv = ob.marching_cubes(D, mn, mx, 0.0)
    n_used = int(rI.max()) + 1
    assert nv == n_used and V.shape == rV.shape and I.size == rI.size
    assert np.all(rV[n_used:] == 0)                                                   # padding to a multiple of 128
    assert np.array_equal(sort_rows(vertex_rows(V[:nv])), sort_rows(vertex_rows(rV[:n_used])))
    assert np.array_equal(canonical_triangles(V, I), canonical_triangles(rV, rI))
    a = vertex_rows(V[:nv]); b = vertex_rows(rV[:n_used])
    oa = np.lexsort(a.T[::-1]); ob_ = np.lexsort(b.T[::-1])
    # pair the vertices through their positions; positions that occur twice (an SDF value exactly on the threshold) are left out
    sa = a[oa]; dup = np.concatenate([[False], np.all(sa[1:] == sa[:-1], axis=1)]); dup = dup | np.concatenate([dup[1:], [False]])
    assert (~dup).mean() > 0.9
    Na, Nb = N[:nv][oa][~dup], rN[:n_used][ob_][~dup]
    assert (np.linalg.norm(Na - Nb, axis=1) / np.maximum(np.linalg.norm(Nb, axis=1), 1e-30)).max() < 1e-5


@pytest.mark.parametrize("ext", ["obj", "ply"])
def test_writer_reproduces_reference_file_bytes(gold, tmp_path, ext):
    w = gold["writer"]
    path = tmp_path / ("m." + ext)
    ob.save_mesh(path, gold["verts"], gold["normals"], gold["colors"], gold["indices"], float(w[0]), tuple(w[1:4]), float(w[4]), tuple(w[5:8]), bool(w[8]))
    assert open(path, "rb").read() == gold[ext].tobytes()
