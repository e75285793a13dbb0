"""Finite-difference pin of the oracle's analytic backward (first AND second order), CPU only.

The reference derives dL/dparams by hand: colour-MLP / SDF-MLP back-propagation plus the double-backward through
the analytic normal n = d sdf / d x (nerf_network.h:376-447; grid.h:556-683,858-883; fully_fused_mlp.cu:1036-1142).
In smooth double precision the restatement must equal the numerical derivative of
    L = sum_i  d[0:3].rgb_i + d[3].sdf_i + (d[4:7]/N + d[8:11]).n_i + d[7].var
"""
import numpy as np
import pytest
from oracle_binding import Oracle


@pytest.mark.parametrize("rgb_hidden,width", [(2, 32), (1, 32), (2, 64)])
def test_backward_matches_finite_differences(rgb_hidden, width):
    o = Oracle(n_levels=4, log2_hashmap=8, base_res=4, top_res=32.0, sdf_width=width, sdf_hidden=1, rgb_width=width, rgb_hidden=rgb_hidden)
    rs = np.random.RandomState(rgb_hidden * 10 + width)
    P = rs.randn(o.n_params) * 0.3
    P[o.off_grid:o.off_var] = rs.randn(o.off_var - o.off_grid) * 0.5
    n, nb, vl = 6, 7, 4
    coords = rs.rand(n, 7).astype(np.float32)
    dout = rs.randn(n, 16); dout[:, 11:] = 0

    def L(P):
        out = o.forward_f64(P, coords, vl)
        return ((dout[:, 0:3] * out[:, 0:3]).sum() + (dout[:, 3] * out[:, 3]).sum() + ((dout[:, 4:7] / nb + dout[:, 8:11]) * out[:, 4:7]).sum() + (dout[:, 7] * out[:, 7]).sum())

    g = o.backward_f64(P, coords, dout, nb, vl)
    touched = o.off_grid + np.nonzero(g[o.off_grid:o.off_var])[0]
    idx = list(rs.choice(o.off_grid, 60, replace=False)) + list(touched[:60]) + [o.off_var]
    assert len(touched) > 20
    for i in idx:
        e = 1e-6
        Pp = P.copy(); Pp[i] += e
        Pm = P.copy(); Pm[i] -= e
        fd = (L(Pp) - L(Pm)) / (2 * e)
        assert abs(fd - g[i]) <= 2e-5 * max(1.0, abs(fd)), (i, fd, g[i])


def test_progressive_levels_zero_the_disabled_features():
    o = Oracle(n_levels=4, log2_hashmap=8, base_res=4, top_res=32.0, sdf_width=32, sdf_hidden=1, rgb_width=32, rgb_hidden=2)
    rs = np.random.RandomState(3)
    P = rs.randn(o.n_params) * 0.3
    coords = rs.rand(4, 7).astype(np.float32)
    dout = rs.randn(4, 16)
    g = o.backward_f64(P, coords, dout, 5, 1)          # valid_level 1 -> levels 0,1 live
    off, _, _ = o.grid_meta()
    assert np.all(g[o.off_grid + 2 * int(off[2]):o.off_var] == 0)
    assert np.any(g[o.off_grid:o.off_grid + 2 * int(off[2])] != 0)
