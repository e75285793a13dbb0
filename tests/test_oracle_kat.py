"""Known-answer tests that pin the oracle's integer / RNG building blocks (CPU only).

The reference ships no golden vectors for this path; these vectors come from public sources
(pcg-random.org demo output) and from the constants the reference's own config implies (SURVEY.md §8 header).
"""
import ctypes as C
import numpy as np
from oracle_binding import Oracle, lib, pcg32, _p
from common import FULL, SMALL


def test_pcg32_matches_published_demo_vector():
    # pcg32_srandom(42, 54): first six outputs printed by the canonical pcg32-demo (pcg-random.org, pcg-c-basic)
    L = lib()
    out = np.zeros(6, np.uint32)
    L.orc_pcg32_seq.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_uint32, C.POINTER(C.c_uint32)]
    L.orc_pcg32_seq(42, 54, 0, 6, _p(out, C.c_uint32))
    assert [hex(int(x)) for x in out] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]


def test_pcg32_advance_equals_stepping():
    u0, _ = pcg32(1337, 0, 40)
    u1, _ = pcg32(1337, 17, 8)
    assert np.array_equal(u0[17:25], u1)
    _, f = pcg32(7, 0, 1000)
    assert f.min() >= 0.0 and f.max() < 1.0


def test_morton_roundtrip_and_known_codes():
    L = lib()
    assert L.orc_morton3D(1, 0, 0) == 1 and L.orc_morton3D(0, 1, 0) == 2 and L.orc_morton3D(0, 0, 1) == 4
    assert L.orc_morton3D(127, 127, 127) == 128 ** 3 - 1
    rs = np.random.RandomState(0)
    for x, y, z in rs.randint(0, 128, (200, 3)):
        m = L.orc_morton3D(int(x), int(y), int(z))
        assert (L.orc_morton3D_invert(m), L.orc_morton3D_invert(m >> 1), L.orc_morton3D_invert(m >> 2)) == (x, y, z)


def test_default_grid_tables_match_reference_config():
    # configs/nerf/base.json -> resolutions / table size quoted in SURVEY.md §8 (derived from grid.h:977-1013)
    o = Oracle(**FULL)
    off, res, sc = o.grid_meta()
    assert list(res) == [16, 24, 34, 50, 72, 104, 151, 219, 317, 461, 669, 971, 1411, 2049]
    assert np.array_equal(sc, res.astype(np.float32) - 1)
    assert int(off[-1]) == 5274064
    assert o.n_params == 10559396 and o.off_rgb == 3072 and o.off_grid == 3072 + 8192
    assert abs(o.per_level_scale - 1.45242) < 1e-4
    s = Oracle(**SMALL)
    assert int(s.grid_meta()[0][-1]) == 118784 and s.sdf_in == 32


def test_grid_index_dense_and_hashed():
    L = lib()
    assert L.orc_grid_index(4096, 16, 3, 2, 1) == (3 + 2 * 16 + 1 * 256) * 2            # dense level
    h = ((5 * 1) ^ ((7 * 2654435761) & 0xFFFFFFFF) ^ ((9 * 805459861) & 0xFFFFFFFF)) % (1 << 19)
    assert L.orc_grid_index(1 << 19, 2049, 5, 7, 9) == h * 2                             # hashed level


def test_progressive_level_schedule():
    o = Oracle(**FULL)
    assert o.valid_level(0) == 14            # quirk: step <= 0 trains all levels (grid.h:1432-1435)
    assert o.valid_level(1) == 3             # ceil(0.2*14) = 3
    assert o.valid_level(100) == 3 and o.valid_level(151) == 4 and o.valid_level(10000) == 14
