"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
from oracle_binding import Oracle, default_flags as orc_flags

SMALL = dict(n_levels=8, log2_hashmap=14, base_res=16, top_res=2048.0, sdf_width=32, sdf_hidden=1, rgb_width=32, rgb_hidden=1)
MID = dict(n_levels=14, log2_hashmap=15, base_res=16, top_res=2048.0, sdf_width=64, sdf_hidden=1, rgb_width=64, rgb_hidden=2)
FULL = dict(n_levels=14, log2_hashmap=19, base_res=16, top_res=2048.0, sdf_width=64, sdf_hidden=1, rgb_width=64, rgb_hidden=2)


def product_config(pkg, o, **kw):
    d = dict(n_levels=o["n_levels"], log2_hashmap_size=o["log2_hashmap"], base_resolution=o["base_res"], top_resolution=o["top_res"],
             sdf_n_neurons=o["sdf_width"], sdf_n_hidden_layers=o["sdf_hidden"], rgb_n_neurons=o["rgb_width"], rgb_n_hidden_layers=o["rgb_hidden"])
    d.update(kw)
    return pkg.default_config(**d)


def copy_flags(pkg, f):
    g = pkg.default_flags()
    for name, _ in f._fields_:
        setattr(g, name, getattr(f, name))
    return g


def random_params(o, seed=0, grid_scale=0.05):
    """Random but well-conditioned parameters: geometric-ish SDF weights so that normals are O(1)."""
    rs = np.random.RandomState(seed)
    p = np.zeros(o.n_params, np.float32)
    p[:o.off_grid] = rs.uniform(-0.3, 0.3, o.off_grid)
    p[o.off_grid:o.off_var] = rs.uniform(-grid_scale, grid_scale, o.off_var - o.off_grid)
    p[o.off_var:] = 0.3
    return p


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_pair(pkg, cfg_dict, views=None, flags=None, threads=4, seed_params=None, **cfg_kw):
    """Oracle + product with identical parameters (taken from the product's reference-order initialisation)."""
    o = Oracle(threads=threads, **cfg_dict)
    t = pkg.Testbed(product_config(pkg, cfg_dict, **cfg_kw))
    if seed_params is None:
        t.init_params()
        p = t.get_params()
    else:
        p = random_params(o, seed_params)
        t.set_params(p)
    o.set_params(p)
    assert o.n_params == t.n_params
    f = flags if flags is not None else orc_flags()
    o.set_flags(f); t.set_flags(copy_flags(pkg, f))
    if views is not None:
        o.set_views(views); t.load_training_data(views)
    return o, t
