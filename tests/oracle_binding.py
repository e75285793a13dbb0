"""ctypes binding of oracle/librnb_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "librnb_oracle.so")


def build_oracle(force=False):
    src = [os.path.join(_ROOT, "oracle", f) for f in ("rnb_oracle.cpp", "orc_common.h", "orc_network.h", "orc_render.h", "orc_mesh.h", "orc_mc_tables.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


class View(C.Structure):
    _fields_ = [("normal_px", C.c_void_p), ("albedo_px", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("xform", C.c_float * 12)]


class Flags(C.Structure):
    _fields_ = [("apply_L2", C.c_int32), ("apply_supernormal", C.c_int32), ("apply_rgbplus", C.c_int32), ("apply_relu", C.c_int32),
                ("apply_bce", C.c_int32), ("light_opti", C.c_int32), ("no_albedo", C.c_int32),
                ("mask_loss_weight", C.c_float), ("ek_loss_weight", C.c_float), ("cos_anneal_ratio", C.c_float), ("light_mode", C.c_int32)]


def default_flags(**kw):
    f = Flags(1, 0, 1, 0, 0, 0, 1, 1.0, 0.01, 1.0, -1)
    for k, v in kw.items():
        setattr(f, k, v)
    return f


class Stats(C.Structure):
    _fields_ = [("loss", C.c_float), ("ek_loss", C.c_float), ("mask_loss", C.c_float), ("n_rays_kept", C.c_uint32),
                ("n_samples", C.c_uint32), ("n_compacted", C.c_uint32), ("n_emitted", C.c_uint32), ("rays_per_batch_next", C.c_uint32)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int]
        L.orc_per_level_scale.restype = C.c_float
        L.orc_per_level_scale.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32]
        L.orc_n_params.restype = C.c_uint64
        L.orc_grid_index.restype = C.c_uint32
        L.orc_valid_level.restype = C.c_uint32
        L.orc_morton3D.restype = C.c_uint32
        L.orc_morton3D_invert.restype = C.c_uint32
        L.orc_compact.restype = C.c_uint32
        L.orc_rollover_weight.restype = C.c_float
        L.orc_density_mean.restype = C.c_float
        L.orc_prep_if_due.restype = C.c_int
        L.orc_get_last_losses.restype = C.c_uint32
        L.orc_marching_cubes.restype = C.c_void_p
        L.orc_save_mesh.restype = C.c_int
        _lib = L
    return _lib


class Oracle:
    """Thin object wrapper.  Shapes follow the reference buffers (coords [n,7], outputs [n,16] ...)."""

    def __init__(self, n_levels=14, log2_hashmap=19, base_res=16, top_res=2048.0, sdf_width=64, sdf_hidden=1, rgb_width=64, rgb_hidden=2,
                 sdf_bias=-0.1, threads=1, per_level_scale=None):
        L = lib()
        self.L = L
        pls = per_level_scale if per_level_scale is not None else L.orc_per_level_scale(C.c_float(top_res), C.c_float(1.0), base_res, n_levels)
        self.per_level_scale = float(pls)
        self.h = C.c_void_p(L.orc_create(n_levels, log2_hashmap, base_res, C.c_float(pls), sdf_width, sdf_hidden, rgb_width, rgb_hidden, C.c_float(sdf_bias), threads))
        self.n_levels = n_levels
        lay = np.zeros(7, np.uint64)
        L.orc_layout(self.h, _p(lay, C.c_uint64))
        self.off_sdf, self.off_rgb, self.off_grid, self.off_var, self.n_params, self.sdf_in, self.rgb_in = [int(x) for x in lay]
        self._keep = []

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def grid_meta(self):
        off = np.zeros(self.n_levels + 1, np.uint32); res = np.zeros(self.n_levels, np.uint32); sc = np.zeros(self.n_levels, np.float32)
        self.L.orc_grid_meta(self.h, _p(off, C.c_uint32), _p(res, C.c_uint32), _p(sc, C.c_float))
        return off, res, sc

    def valid_level(self, step):
        return int(self.L.orc_valid_level(self.h, int(step)))

    def init_params(self, seed=1337, sdf_init=None):
        if sdf_init is not None:
            sdf_init = np.ascontiguousarray(sdf_init, np.float32)
            self.L.orc_init_params(self.h, C.c_uint32(seed), _p(sdf_init, C.c_float), C.c_uint64(sdf_init.size))
        else:
            self.L.orc_init_params(self.h, C.c_uint32(seed), None, C.c_uint64(0))

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32); assert p.size == self.n_params
        self.L.orc_set_params(self.h, _p(p, C.c_float))

    def get_params(self):
        m = np.zeros(self.n_params, np.float32); hv = np.zeros_like(m); e = np.zeros_like(m)
        self.L.orc_get_params(self.h, _p(m, C.c_float), _p(hv, C.c_float), _p(e, C.c_float))
        return m, hv, e

    def get_grads(self):
        g = np.zeros(self.n_params, np.float32)
        self.L.orc_get_grads(self.h, _p(g, C.c_float)); return g

    def set_grads(self, g):
        g = np.ascontiguousarray(g, np.float32); self.L.orc_set_grads(self.h, _p(g, C.c_float))

    def get_opt_state(self):
        m1 = np.zeros(self.n_params, np.float32); m2 = np.zeros_like(m1); st = np.zeros(self.n_params, np.uint32)
        self.L.orc_get_opt_state(self.h, _p(m1, C.c_float), _p(m2, C.c_float), _p(st, C.c_uint32)); return m1, m2, st

    def set_views(self, views):
        """views: list of dicts(normal=uint16[h,w,4], albedo=uint16[h,w,4]|None, fx, fy, cx, cy, xform=float32[12] col-major)"""
        arr = (View * len(views))()
        self._keep = []
        for i, v in enumerate(views):
            n = np.ascontiguousarray(v["normal"], np.uint16); self._keep.append(n)
            arr[i].normal_px = n.ctypes.data
            if v.get("albedo") is not None:
                a = np.ascontiguousarray(v["albedo"], np.uint16); self._keep.append(a); arr[i].albedo_px = a.ctypes.data
            else:
                arr[i].albedo_px = None
            arr[i].h, arr[i].w = n.shape[0], n.shape[1]
            arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = v["fx"], v["fy"], v["cx"], v["cy"]
            for k in range(12):
                arr[i].xform[k] = float(v["xform"][k])
        self._views = arr
        self.L.orc_set_views(self.h, arr, len(views))

    def set_flags(self, flags):
        self.flags = flags; self.L.orc_set_flags(self.h, C.byref(flags))

    def set_train_state(self, training_step=0, rays_per_batch=4096, n_rays_total=0, measured_before=0, pin_rays=1, target_batch=1 << 18):
        self.L.orc_set_train_state(self.h, training_step, rays_per_batch, n_rays_total, measured_before, pin_rays, target_batch)

    def set_canonical_state(self, canonical_step, n_images_prev):
        self.L.orc_set_canonical_state(self.h, canonical_step, n_images_prev)

    def get_rng(self):
        o = np.zeros(4, np.uint64); self.L.orc_get_rng(self.h, _p(o, C.c_uint64)); return [int(x) for x in o]

    def set_rng(self, s, i, ds, di):
        self.L.orc_set_rng(self.h, C.c_uint64(s), C.c_uint64(i), C.c_uint64(ds), C.c_uint64(di))

    def get_bitfield(self):
        b = np.zeros(128 ** 3 * 8 // 8, np.uint8); self.L.orc_get_bitfield(self.h, _p(b, C.c_uint8)); return b

    def set_bitfield(self, b):
        b = np.ascontiguousarray(b, np.uint8); assert b.size == 128 ** 3; self.L.orc_set_bitfield(self.h, _p(b, C.c_uint8))

    def get_density_grid(self):
        g = np.zeros(128 ** 3, np.float32); self.L.orc_get_density_grid(self.h, _p(g, C.c_float)); return g

    def set_density_grid(self, g, ema_step):
        g = np.ascontiguousarray(g, np.float32); assert g.size == 128 ** 3; self.L.orc_set_density_grid(self.h, _p(g, C.c_float), ema_step)

    def generate_samples(self, n_rays, n_rays_total, max_samples):
        ri = np.zeros(n_rays, np.uint32); rays = np.zeros((n_rays, 6), np.float32); ns = np.zeros((n_rays, 2), np.uint32)
        coords = np.zeros((max_samples, 7), np.float32); cnt = np.zeros(2, np.uint32)
        self.L.orc_generate_samples(self.h, n_rays, n_rays_total, max_samples, _p(ri, C.c_uint32), _p(rays, C.c_float), _p(ns, C.c_uint32), _p(coords, C.c_float), _p(cnt, C.c_uint32))
        k = int(cnt[0])
        return dict(ray_indices=ri[:k], rays=rays[:k], numsteps=ns[:k], coords=coords, n_kept=k, n_samples=int(cnt[1]))

    def network_forward(self, coords, valid_level, use_ema=False, with_rgb=True):
        coords = np.ascontiguousarray(coords, np.float32); n = coords.shape[0]
        out = np.zeros((n, 16), np.float32); nrm = np.zeros((n, 3), np.float32)
        self.L.orc_network_forward(self.h, _p(coords, C.c_float), C.c_uint64(n), valid_level, int(use_ema), int(with_rgb), _p(out, C.c_float), _p(nrm, C.c_float))
        return out, nrm

    def encode(self, xyz, valid_level):
        xyz = np.ascontiguousarray(xyz, np.float32); n = xyz.shape[0]
        enc = np.zeros((n, 2 * self.n_levels), np.float32); dydx = np.zeros((n, 2 * self.n_levels, 3), np.float32)
        self.L.orc_encode(self.h, _p(xyz, C.c_float), C.c_uint64(n), valid_level, _p(enc, C.c_float), _p(dydx, C.c_float))
        return enc, dydx

    def eval_sdf(self, xyz, valid_level, use_ema=False):
        xyz = np.ascontiguousarray(xyz, np.float32); n = xyz.shape[0]
        s = np.zeros(n, np.float32); d = np.zeros(n, np.float32)
        self.L.orc_eval_sdf(self.h, _p(xyz, C.c_float), C.c_uint64(n), valid_level, int(use_ema), _p(s, C.c_float), _p(d, C.c_float))
        return s, d

    def compact(self, out_a, numsteps, max_compacted):
        out_a = np.ascontiguousarray(out_a, np.float32); numsteps = np.ascontiguousarray(numsteps, np.uint32); k = numsteps.shape[0]
        nf = np.zeros(k, np.uint32); cb = np.zeros(k, np.uint32); ne = np.zeros(k, np.uint32)
        tot = self.L.orc_compact(self.h, _p(out_a, C.c_float), _p(numsteps, C.c_uint32), k, max_compacted, _p(nf, C.c_uint32), _p(cb, C.c_uint32), _p(ne, C.c_uint32))
        return nf, cb, ne, int(tot)

    def loss(self, out_c, ray_indices, n_fwd, cbase, n_emit, n_rays, n_rays_total, step):
        out_c = np.ascontiguousarray(out_c, np.float32); k = len(ray_indices)
        dout = np.zeros_like(out_c); lo = np.zeros(k, np.float32); ek = np.zeros(k, np.float32); ml = np.zeros(k, np.float32)
        ri = np.ascontiguousarray(ray_indices, np.uint32)
        self.L.orc_loss(self.h, _p(out_c, C.c_float), _p(ri, C.c_uint32), _p(np.ascontiguousarray(n_fwd, np.uint32), C.c_uint32), _p(np.ascontiguousarray(cbase, np.uint32), C.c_uint32),
                        _p(np.ascontiguousarray(n_emit, np.uint32), C.c_uint32), k, n_rays, n_rays_total, step, _p(dout, C.c_float), _p(lo, C.c_float), _p(ek, C.c_float), _p(ml, C.c_float))
        return dout, lo, ek, ml

    def network_backward(self, coords, dout, n_in, n_batch, valid_level):
        coords = np.ascontiguousarray(coords, np.float32); dout = np.ascontiguousarray(dout, np.float32)
        self.L.orc_network_backward(self.h, _p(coords, C.c_float), _p(dout, C.c_float), C.c_uint64(coords.shape[0]), n_in, n_batch, valid_level)
        return self.get_grads()

    def optimizer_step(self):
        self.L.orc_optimizer_step(self.h)

    def density_update(self, n_uniform, n_nonuniform, valid_level):
        self.L.orc_density_update(self.h, n_uniform, n_nonuniform, valid_level)

    def set_world(self, world, rank):
        self.L.orc_set_world(self.h, world, rank)

    def set_dp_exact(self, on):
        self.L.orc_set_dp_exact(self.h, int(on))

    def set_opt_shard(self, begin, end):
        self.L.orc_set_opt_shard(self.h, C.c_uint64(begin), C.c_uint64(end))

    def set_half_params(self, hv):
        hv = np.ascontiguousarray(hv, np.float32); assert hv.size == self.n_params
        self.L.orc_set_half_params(self.h, _p(hv, C.c_float))

    def train_step_begin(self):
        self.L.orc_train_step_begin(self.h)

    def train_step_end(self):
        st = Stats(); self.L.orc_train_step_end(self.h, C.byref(st)); return st

    def get_sums(self):
        a = np.zeros(4, np.float64); self.L.orc_get_sums(self.h, _p(a, C.c_double)); return a

    def set_sums(self, a):
        a = np.ascontiguousarray(a, np.float64); self.L.orc_set_sums(self.h, _p(a, C.c_double))

    def train_step(self):
        st = Stats(); self.L.orc_train_step(self.h, C.byref(st)); return st

    def last_losses(self, cap=1 << 18):
        ri = np.zeros(cap, np.uint32); lo = np.zeros(cap, np.float32); ek = np.zeros(cap, np.float32); ml = np.zeros(cap, np.float32)
        k = int(self.L.orc_get_last_losses(self.h, cap, _p(ri, C.c_uint32), _p(lo, C.c_float), _p(ek, C.c_float), _p(ml, C.c_float)))
        return ri[:k], lo[:k], ek[:k], ml[:k]

    def forward_f64(self, params, coords, valid_level):
        params = np.ascontiguousarray(params, np.float64); coords = np.ascontiguousarray(coords, np.float32); n = coords.shape[0]
        out = np.zeros((n, 16), np.float64)
        self.L.orc_forward_f64(self.h, _p(params, C.c_double), _p(coords, C.c_float), C.c_uint64(n), valid_level, _p(out, C.c_double)); return out

    def backward_f64(self, params, coords, dout, n_batch, valid_level):
        params = np.ascontiguousarray(params, np.float64); coords = np.ascontiguousarray(coords, np.float32); dout = np.ascontiguousarray(dout, np.float64)
        g = np.zeros(self.n_params, np.float64)
        self.L.orc_backward_f64(self.h, _p(params, C.c_double), _p(coords, C.c_float), _p(dout, C.c_double), C.c_uint64(coords.shape[0]), n_batch, valid_level, _p(g, C.c_double)); return g


def pcg32(seed, advance, n):
    u = np.zeros(n, np.uint32); f = np.zeros(n, np.float32)
    lib().orc_pcg32(C.c_uint64(seed), C.c_int64(advance), n, _p(u, C.c_uint32), _p(f, C.c_float)); return u, f


# ---- mesh path (oracle/orc_mesh.h) ----
def marching_cubes(density, aabb_min=(0, 0, 0), aabb_max=(1, 1, 1), thresh=0.0):
    """density[z, y, x] float32 -> (verts [Vp,3] incl. zero padding to a multiple of 128, normals [Vp,3], indices [T*3], n_verts)."""
    d = np.ascontiguousarray(density, np.float32)
    res = (C.c_uint32 * 3)(d.shape[2], d.shape[1], d.shape[0])
    mn = (C.c_float * 3)(*aabb_min); mx = (C.c_float * 3)(*aabb_max); cnt = (C.c_uint32 * 3)()
    L = lib()
    h = C.c_void_p(L.orc_marching_cubes(_p(d, C.c_float), res, mn, mx, C.c_float(thresh), cnt))
    verts = np.zeros((cnt[1], 3), np.float32); normals = np.zeros((cnt[1], 3), np.float32); idx = np.zeros(cnt[2], np.uint32)
    L.orc_mesh_get(h, _p(verts, C.c_float), _p(normals, C.c_float), _p(idx, C.c_uint32))
    L.orc_mesh_free(h)
    return verts, normals, idx, int(cnt[0])


def save_mesh(path, verts, normals, colors, indices, nerf_scale=1.0, nerf_offset=(0, 0, 0), n2w_s=1.0, n2w_t=(0, 0, 0), invert_normals=False):
    v = np.ascontiguousarray(verts, np.float32); n = np.ascontiguousarray(normals, np.float32); c = np.ascontiguousarray(colors, np.float32)
    i = np.ascontiguousarray(indices, np.uint32)
    rc = lib().orc_save_mesh(_p(v, C.c_float), _p(n, C.c_float), _p(c, C.c_float), _p(i, C.c_uint32), C.c_uint32(v.shape[0]), C.c_uint32(i.size), str(path).encode(),
                             C.c_float(nerf_scale), (C.c_float * 3)(*nerf_offset), C.c_float(n2w_s), (C.c_float * 3)(*n2w_t), int(invert_normals))
    assert rc == 0, "oracle save_mesh failed"
