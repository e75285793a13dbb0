"""rnb-neus2_b200 — host-side mirror of the reference Testbed training interface over the C ABI of librnb_b200.so.

The product is the CUDA library (csrc/, include/rnb_b200.h).  This module is the thin Python host layer used by
tests/ and bench.py: it mirrors the names of the reference's C++ host interface for this path
(Testbed::train / training_prep_nerf / train_nerf, Trainer::optimizer_step, NerfNetwork::sdf — reference
src/testbed.cu:2776-2872, src/testbed_nerf.cu:3560-3668,4125-4138).

The directory name contains a hyphen, so import it with `load_package()` from `rnb_loader.py` (repo root) or
importlib.  There is NO CPU fallback: if the CUDA library is missing or no GPU is present, calls raise.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "librnb_b200.so")
GRID_CELLS = 128 ** 3


class RnbError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_levels", C.c_uint32), ("log2_hashmap_size", C.c_uint32), ("base_resolution", C.c_uint32),
                ("per_level_scale", C.c_float), ("top_resolution", C.c_float), ("base_valid_level_scale", C.c_float), ("valid_level_scale", C.c_float),
                ("base_training_step", C.c_uint32), ("sdf_n_neurons", C.c_uint32), ("sdf_n_hidden_layers", C.c_uint32), ("rgb_n_neurons", C.c_uint32),
                ("rgb_n_hidden_layers", C.c_uint32), ("sdf_bias", C.c_float), ("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("epsilon", C.c_float), ("l2_reg", C.c_float), ("ema_decay", C.c_float), ("lr_decay_start", C.c_uint32), ("lr_decay_interval", C.c_uint32),
                ("lr_decay_base", C.c_float), ("loss_scale", C.c_float), ("target_batch_size", C.c_uint32), ("rays_per_batch", C.c_uint32),
                ("pin_rays_per_batch", C.c_uint32), ("seed", C.c_uint32), ("density_grid_decay", C.c_float), ("world_size", C.c_uint32), ("rank", C.c_uint32)]


class Flags(C.Structure):
    _fields_ = [("apply_L2", C.c_int32), ("apply_supernormal", C.c_int32), ("apply_rgbplus", C.c_int32), ("apply_relu", C.c_int32), ("apply_bce", C.c_int32),
                ("light_opti", C.c_int32), ("no_albedo", C.c_int32), ("mask_loss_weight", C.c_float), ("ek_loss_weight", C.c_float), ("cos_anneal_ratio", C.c_float),
                ("light_mode", C.c_int32), ("only_sdf_training", C.c_int32)]


class View(C.Structure):
    _fields_ = [("normal_px", C.c_void_p), ("albedo_px", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("xform", C.c_float * 12)]


class StepStats(C.Structure):
    _fields_ = [("loss", C.c_float), ("ek_loss", C.c_float), ("mask_loss", C.c_float), ("n_rays", C.c_uint32), ("n_rays_kept", C.c_uint32),
                ("n_samples", C.c_uint32), ("n_samples_compacted", C.c_uint32), ("n_samples_trained", C.c_uint32), ("rays_per_batch_next", C.c_uint32),
                ("training_step", C.c_uint32), ("density_grid_updated", C.c_uint32)]


EXPORTED_SYMBOLS = [
    "rnb_last_error", "rnb_abi_version", "rnb_create", "rnb_destroy", "rnb_default_config", "rnb_default_flags", "rnb_param_layout", "rnb_init_params",
    "rnb_set_params_fp32", "rnb_get_params_fp32", "rnb_export_params_fp16", "rnb_import_params_fp16", "rnb_export_density_grid", "rnb_import_density_grid",
    "rnb_get_bitfield", "rnb_set_bitfield", "rnb_get_train_state", "rnb_set_train_state", "rnb_set_canonical_state", "rnb_get_rng", "rnb_set_rng", "rnb_set_dataset", "rnb_upload_dataset",
    "rnb_set_flags", "rnb_prep", "rnb_train_step", "rnb_train", "rnb_train_step_begin", "rnb_train_step_end", "rnb_grad_buffer", "rnb_comm_unique_id", "rnb_comm_init", "rnb_comm_adopt", "rnb_comm_destroy", "rnb_comm_info", "rnb_comm_sync_ema", "rnb_stat_buffer", "rnb_param_buffers", "rnb_set_optimizer_shard", "rnb_get_grads_fp32", "rnb_get_ray_losses", "rnb_get_ray_counts", "rnb_checkpoint_save", "rnb_checkpoint_restore", "rnb_profile_enable", "rnb_profile_read", "rnb_launch_count", "rnb_eval_sdf", "rnb_sdf_on_grid",
    "rnb_load_png_rgba16", "rnb_free_host", "rnb_load_dataset_images", "rnb_marching_cubes", "rnb_marching_cubes_from_density", "rnb_mesh_buffers", "rnb_mesh_download", "rnb_save_mesh",
    "rnb_stage_generate", "rnb_stage_forward", "rnb_stage_loss", "rnb_stage_backward", "rnb_stage_optimizer",
    "rnb_raymesh_create", "rnb_raymesh_destroy", "rnb_raymesh_info", "rnb_raymesh_intersect",
]


class MeshInfo(C.Structure):
    _fields_ = [("n_verts", C.c_uint32), ("n_verts_padded", C.c_uint32), ("n_indices", C.c_uint32), ("res", C.c_uint32 * 3), ("stage_ms", C.c_float * 4)]


def save_mesh_device(path, verts_ptr, normals_ptr, colors_ptr, indices_ptr, n_verts, n_indices, nerf_scale=1.0, nerf_offset=(0.0, 0.0, 0.0), n2w_s=1.0, n2w_t=(0.0, 0.0, 0.0),
                     invert_normals=False, stream=None):
    """rnb_save_mesh on device arrays (OBJ, or ASCII PLY for a path ending in "ply"); returns the bytes written."""
    nb = C.c_uint64(0)
    rc = lib().rnb_save_mesh(C.c_void_p(verts_ptr), C.c_void_p(normals_ptr), C.c_void_p(colors_ptr), C.c_void_p(indices_ptr), C.c_uint32(n_verts), C.c_uint32(n_indices), str(path).encode(),
                             C.c_float(nerf_scale), (C.c_float * 3)(*[float(x) for x in nerf_offset]), C.c_float(n2w_s), (C.c_float * 3)(*[float(x) for x in n2w_t]), int(invert_normals),
                             C.c_void_p(stream), C.byref(nb))
    if rc != 0:
        raise RnbError(lib().rnb_last_error().decode())
    return int(nb.value)


class RayMesh:
    """Ray queries against a triangle mesh on the GPU (rnb_raymesh_*, csrc/rnb_raymesh.cu) — what the reference's albedo-scaling stage
    asks of trimesh (`mesh.ray.intersects_location`, rnb_neus2/albedo_scaling.py:285-289,316-329).  No CPU path."""
    NO_TRI = 0xFFFFFFFF

    def __init__(self, verts, tris, grid_res=0):
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        self.tris = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
        self.h = C.c_void_p()
        L = lib()
        rc = L.rnb_raymesh_create(_p(self.verts, C.c_float), C.c_uint32(self.verts.shape[0]), _p(self.tris, C.c_uint32), C.c_uint32(self.tris.shape[0]), C.c_uint32(grid_res), C.byref(self.h))
        if rc != 0:
            raise RnbError(L.rnb_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            lib().rnb_raymesh_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        res = (C.c_uint32 * 3)(); refs = C.c_uint64(0)
        if lib().rnb_raymesh_info(self.h, res, C.byref(refs)) != 0:
            raise RnbError(lib().rnb_last_error().decode())
        return dict(res=tuple(int(x) for x in res), n_cell_refs=int(refs.value))

    def _run(self, origins, dirs, t_max):
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3); d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("origins and directions differ in shape")
        n = o.shape[0]
        tm = np.ascontiguousarray(t_max, np.float64).reshape(-1) if t_max is not None else None
        if tm is not None and tm.shape[0] != n:
            raise ValueError("t_max needs one entry per ray")
        t = np.full(n, np.inf, np.float64); tri = np.full(n, self.NO_TRI, np.uint32)
        rc = lib().rnb_raymesh_intersect(self.h, _p(o, C.c_double), _p(d, C.c_double), _p(tm, C.c_double), C.c_uint32(n), _p(t, C.c_double), _p(tri, C.c_uint32), C.c_void_p(None))
        if rc != 0:
            raise RnbError(lib().rnb_last_error().decode())
        return t, tri

    def first_hit(self, origins, dirs):
        """closest intersection with t > 0 per ray: (t, triangle); t = inf and triangle = NO_TRI for a miss"""
        return self._run(origins, dirs, None)

    def any_hit(self, origins, dirs, t_max):
        """bool per ray: some intersection with 0 < t < t_max"""
        return self._run(origins, dirs, t_max)[1] != self.NO_TRI


def load_png_rgba16(path):
    """rnb_load_png_rgba16: a PNG as stbi_load_16(..., 4) returns it — uint16 [h, w, 4].  Host code (works without a GPU)."""
    w = C.c_uint32(); h = C.c_uint32(); px = C.POINTER(C.c_uint16)()
    L = lib()
    if L.rnb_load_png_rgba16(str(path).encode(), C.byref(w), C.byref(h), C.byref(px)) != 0:
        raise RnbError(L.rnb_last_error().decode())
    try:
        return np.ctypeslib.as_array(px, shape=(h.value, w.value, 4)).copy()
    finally:
        L.rnb_free_host(px)


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into librnb_b200.so (in-tree)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_DIR, "csrc"), "-j8"], stdout=out)
    return _SO


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RnbError("librnb_b200.so is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(_SO)
        L.rnb_last_error.restype = C.c_char_p
        L.rnb_free_host.restype = None
        L.rnb_abi_version.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def default_config(**kw):
    c = Config()
    lib().rnb_default_config(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def default_flags(**kw):
    f = Flags()
    lib().rnb_default_flags(C.byref(f))
    for k, v in kw.items():
        setattr(f, k, v)
    return f


def small_config(**kw):
    """BASELINE configs[0]: L=8, T=2^14, base 16 -> top 2048, 32-wide MLPs with one hidden layer each."""
    base = dict(n_levels=8, log2_hashmap_size=14, base_resolution=16, top_resolution=2048.0, sdf_n_neurons=32, rgb_n_neurons=32, rgb_n_hidden_layers=1)
    base.update(kw)
    return default_config(**base)


class Testbed:
    """Mirror of the reference Testbed's training surface for the NeuS2 step."""

    def __init__(self, config=None, flags=None):
        self.L = lib()
        self.cfg = config if config is not None else default_config()
        self.h = C.c_void_p()
        self._chk(self.L.rnb_create(C.byref(self.cfg), C.byref(self.h)))
        lay = (C.c_uint64 * 5)()
        self._chk(self.L.rnb_param_layout(self.h, lay))
        self.off_sdf, self.off_rgb, self.off_grid, self.off_var, self.n_params = [int(x) for x in lay]
        self._keep = []
        self.last_measured_batch_size = 0; self.last_loss = 0.0       # of the last step whose statistics were read (snapshot fields)
        self._n_images_prev = 0; self._n_views = 0                    # Training::n_images_for_training_prev of this "process" (testbed.h:578)
        if flags is not None:
            self.set_flags(flags)

    def _chk(self, rc):
        if rc != 0:
            raise RnbError("rnb error %d: %s" % (rc, self.L.rnb_last_error().decode()))

    def close(self):
        if self.h:
            self.L.rnb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- parameters / snapshot (Trainer::initialize_params, serialize, deserialize) ---
    def init_params(self, sdf_init=None):
        if sdf_init is None:
            self._chk(self.L.rnb_init_params(self.h, None, C.c_size_t(0)))
        else:
            a = np.ascontiguousarray(sdf_init, np.float32)
            self._chk(self.L.rnb_init_params(self.h, _p(a, C.c_float), C.c_size_t(a.size)))

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32)
        self._chk(self.L.rnb_set_params_fp32(self.h, _p(p, C.c_float), C.c_size_t(p.size)))

    def get_params(self):
        p = np.zeros(self.n_params, np.float32)
        self._chk(self.L.rnb_get_params_fp32(self.h, _p(p, C.c_float), C.c_size_t(p.size)))
        return p

    def export_params_fp16(self, use_ema=False):
        p = np.zeros(self.n_params, np.uint16)
        self._chk(self.L.rnb_export_params_fp16(self.h, _p(p, C.c_uint16), C.c_size_t(p.size), int(use_ema)))
        return p.view(np.float16)

    def import_params_fp16(self, p):
        p = np.ascontiguousarray(p, np.float16).view(np.uint16)
        self._chk(self.L.rnb_import_params_fp16(self.h, _p(p, C.c_uint16), C.c_size_t(p.size)))

    def export_density_grid(self):
        g = np.zeros(GRID_CELLS, np.float32); st = C.c_uint32()
        self._chk(self.L.rnb_export_density_grid(self.h, _p(g, C.c_float), C.c_size_t(g.size), C.byref(st)))
        return g, int(st.value)

    def import_density_grid(self, g, ema_step):
        g = np.ascontiguousarray(g, np.float32)
        self._chk(self.L.rnb_import_density_grid(self.h, _p(g, C.c_float), C.c_size_t(g.size), ema_step))

    def get_bitfield(self):
        b = np.zeros(GRID_CELLS, np.uint8)
        self._chk(self.L.rnb_get_bitfield(self.h, _p(b, C.c_uint8), C.c_size_t(b.size)))
        return b

    def set_bitfield(self, b):
        b = np.ascontiguousarray(b, np.uint8)
        self._chk(self.L.rnb_set_bitfield(self.h, _p(b, C.c_uint8), C.c_size_t(b.size)))

    def get_train_state(self):
        o = (C.c_uint32 * 4)(); self._chk(self.L.rnb_get_train_state(self.h, o)); return [int(x) for x in o]

    def set_train_state(self, training_step, rays_per_batch, n_rays_total=0, measured_before=0):
        self._chk(self.L.rnb_set_train_state(self.h, training_step, rays_per_batch, n_rays_total, measured_before))

    def set_canonical_state(self, canonical_training_step, n_images_prev):
        """The two members Testbed::load_snapshot does not restore: m_canonical_training_step and n_images_for_training_prev (rnb_b200.h)."""
        self._chk(self.L.rnb_set_canonical_state(self.h, canonical_training_step, n_images_prev))

    def get_rng(self):
        o = (C.c_uint64 * 4)(); self._chk(self.L.rnb_get_rng(self.h, o)); return [int(x) for x in o]

    def set_rng(self, vals):
        a = (C.c_uint64 * 4)(*vals); self._chk(self.L.rnb_set_rng(self.h, a))

    # --- dataset (Testbed::load_training_data result) ---
    def _views(self, views, device_ptrs):
        arr = (View * len(views))()
        self._keep = []
        for i, v in enumerate(views):
            if device_ptrs:
                arr[i].normal_px = int(v["normal"]); arr[i].albedo_px = int(v["albedo"]) if v.get("albedo") else None
                arr[i].w, arr[i].h = v["w"], v["h"]
            else:
                n = np.ascontiguousarray(v["normal"], np.uint16); self._keep.append(n)
                arr[i].normal_px = n.ctypes.data
                if v.get("albedo") is not None:
                    a = np.ascontiguousarray(v["albedo"], np.uint16); self._keep.append(a); arr[i].albedo_px = a.ctypes.data
                arr[i].h, arr[i].w = n.shape[0], n.shape[1]
            arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = v["fx"], v["fy"], v["cx"], v["cy"]
            for k in range(12):
                arr[i].xform[k] = float(v["xform"][k])
        return arr

    def load_training_data(self, views):
        """views: list of dict(normal=uint16[h,w,4], albedo=uint16[h,w,4]|None, fx, fy, cx, cy, xform[12]) in host memory."""
        arr = self._views(views, False)
        self._chk(self.L.rnb_upload_dataset(self.h, arr, len(views)))
        self._n_views = len(views)
        self._keep = []

    def set_dataset_device(self, views):
        arr = self._views(views, True)
        self._chk(self.L.rnb_set_dataset(self.h, arr, len(views)))
        self._n_views = len(views)

    def set_flags(self, flags):
        self.flags = flags
        self._chk(self.L.rnb_set_flags(self.h, C.byref(flags)))

    # --- training (Testbed::train / training_prep_nerf / train_nerf) ---
    def training_prep_nerf(self, stream=None):
        self._chk(self.L.rnb_prep(self.h, C.c_void_p(stream)))
        self._n_images_prev = self._n_views

    def train_nerf(self, stream=None, want_stats=True):
        st = StepStats()
        self._chk(self.L.rnb_train_step(self.h, C.c_void_p(stream), C.byref(st) if want_stats else None))
        return st

    def train(self, stream=None, want_stats=True):
        st = StepStats()
        self._chk(self.L.rnb_train(self.h, C.c_void_p(stream), C.byref(st) if want_stats else None))
        self._n_images_prev = self._n_views          # the first Testbed::train of a process always refreshes the occupancy grid (canonical step 0)
        if want_stats:
            self.last_measured_batch_size = int(st.n_samples_compacted); self.last_loss = float(st.loss)
        return st

    def train_step_begin(self, stream=None):
        self._chk(self.L.rnb_train_step_begin(self.h, C.c_void_p(stream)))

    def train_step_end(self, stream=None):
        st = StepStats()
        self._chk(self.L.rnb_train_step_end(self.h, C.c_void_p(stream), C.byref(st)))
        self.last_measured_batch_size = int(st.n_samples_compacted); self.last_loss = float(st.loss)
        return st

    # --- data parallelism behind the boundary (rnb_comm_*) ---
    @staticmethod
    def comm_unique_id():
        """ncclGetUniqueId through the library (call on ONE rank and carry the 128 bytes to the others)."""
        buf = (C.c_uint8 * 128)()
        rc = lib().rnb_comm_unique_id(buf)
        if rc != 0:
            raise RnbError(lib().rnb_last_error().decode())
        return bytes(buf)

    def comm_init(self, unique_id):
        """ncclCommInitRank(world_size, id, rank) on the current device; afterwards train() / train_nerf() exchange the gradients themselves."""
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._chk(self.L.rnb_comm_init(self.h, buf))

    def comm_destroy(self):
        self._chk(self.L.rnb_comm_destroy(self.h))

    def comm_info(self):
        o = (C.c_uint32 * 4)(); self._chk(self.L.rnb_comm_info(self.h, o))
        return dict(installed=bool(o[0]), nccl_version=int(o[1]), sharded=bool(o[2] & 1), one_sample_order=bool(o[2] & 2), world_size=int(o[3]))

    def grad_buffer(self):
        p = C.POINTER(C.c_float)(); n = C.c_uint64()
        self._chk(self.L.rnb_grad_buffer(self.h, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, int(n.value)

    def param_buffers(self):
        """device pointers of the binary16 training / EMA parameters, the parameter count and the padded (allocated, zero-tailed) count"""
        pp = C.c_void_p(); pe = C.c_void_p(); n = C.c_uint64(); npad = C.c_uint64()
        self._chk(self.L.rnb_param_buffers(self.h, C.byref(pp), C.byref(pe), C.byref(n), C.byref(npad)))
        return pp.value, pe.value, int(n.value), int(npad.value)

    def set_optimizer_shard(self, begin, end, reduced_grads_ptr=None):
        """data-parallel optimizer sharding: Adam/EMA on [begin, end) only (rnb_set_optimizer_shard); end == 0 switches it off"""
        self._chk(self.L.rnb_set_optimizer_shard(self.h, C.c_uint64(begin), C.c_uint64(end), C.c_void_p(reduced_grads_ptr)))

    def get_grads(self):
        g = np.zeros(self.n_params, np.float32)
        self._chk(self.L.rnb_get_grads_fp32(self.h, _p(g, C.c_float), C.c_size_t(g.size)))
        return g

    def ray_losses(self, cap=1 << 18):
        ri = np.zeros(cap, np.uint32); l3 = np.zeros((cap, 3), np.float32); n = C.c_uint32(0)
        self._chk(self.L.rnb_get_ray_losses(self.h, C.c_uint32(cap), _p(ri, C.c_uint32), _p(l3, C.c_float), C.byref(n)))
        return ri[:n.value], l3[:n.value]

    def ray_counts(self, cap=1 << 18):
        a = np.zeros(cap, np.uint32); b = np.zeros(cap, np.uint32); n = C.c_uint32(0)
        self._chk(self.L.rnb_get_ray_counts(self.h, C.c_uint32(cap), _p(a, C.c_uint32), _p(b, C.c_uint32), C.byref(n)))
        return a[:n.value], b[:n.value]

    def stat_buffer(self):
        p = C.POINTER(C.c_float)(); n = C.c_uint64()
        self._chk(self.L.rnb_stat_buffer(self.h, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, int(n.value)

    def checkpoint_save(self):
        self._chk(self.L.rnb_checkpoint_save(self.h))

    def checkpoint_restore(self):
        self._chk(self.L.rnb_checkpoint_restore(self.h))

    def profile_enable(self, on=True):
        self._chk(self.L.rnb_profile_enable(self.h, int(on)))

    def profile_read(self):
        buf = C.create_string_buffer(1024); ms = (C.c_double * 32)(); calls = (C.c_uint64 * 32)(); n = C.c_uint32(32)
        self._chk(self.L.rnb_profile_read(self.h, buf, C.c_size_t(1024), ms, calls, C.byref(n)))
        names = buf.value.decode().split(";")[:n.value]
        return {names[i]: (float(ms[i]), int(calls[i])) for i in range(n.value)}

    def launch_count(self):
        v = C.c_uint64(); self._chk(self.L.rnb_launch_count(self.h, C.byref(v))); return int(v.value)

    def eval_sdf_device(self, xyz_dev_ptr, n, sdf_ptr=None, normal_ptr=None, density_ptr=None, use_ema=True, stream=None):
        self._chk(self.L.rnb_eval_sdf(self.h, C.c_void_p(xyz_dev_ptr), C.c_size_t(n), C.c_void_p(sdf_ptr), C.c_void_p(normal_ptr), C.c_void_p(density_ptr), int(use_ema), C.c_void_p(stream)))

    def sdf_on_grid_device(self, res, aabb_min, aabb_max, out_ptr, use_ema=True, stream=None):
        r = (C.c_uint32 * 3)(*[int(x) for x in res]); a = (C.c_float * 3)(*[float(x) for x in aabb_min]); b = (C.c_float * 3)(*[float(x) for x in aabb_max])
        self._chk(self.L.rnb_sdf_on_grid(self.h, r, a, b, C.c_void_p(out_ptr), int(use_ema), C.c_void_p(stream)))

    # --- dataset ingest ---
    def load_training_data_dir(self, path, threads=0, stream=None):
        """Testbed::load_training_data on a scene directory / transform.json (src/testbed.cu load_nerf -> src/nerf_loader.cu): the JSON is read
        on the host (dataset.py), the PNGs are decoded by the library's host threads and uploaded.  Returns the dataset description
        (scale, offset, n2w, from_na, ... — what save_mesh needs later)."""
        from . import dataset as ds
        meta = ds.load_transforms(path)
        n = len(meta["views"])
        arr = (View * n)()
        for i, v in enumerate(meta["views"]):
            arr[i].normal_px = None; arr[i].albedo_px = None; arr[i].w = v["w"]; arr[i].h = v["h"]
            arr[i].fx = float(v["fx"]); arr[i].fy = float(v["fy"]); arr[i].cx = float(v["cx"]); arr[i].cy = float(v["cy"])
            for k in range(12):
                arr[i].xform[k] = float(v["xform"][k])
        normals = (C.c_char_p * n)(*[v["normal_path"].encode() for v in meta["views"]])
        albedos = (C.c_char_p * n)(*[(v["albedo_path"].encode() if v["albedo_path"] else None) for v in meta["views"]])
        self._chk(self.L.rnb_load_dataset_images(self.h, arr, n, normals, albedos, int(threads), C.c_void_p(stream)))
        self._n_views = n
        self.dataset = meta
        return meta

    # --- snapshot hand-off ---
    def save_snapshot(self, path, network_config, loss=None, aabb_scale=1, dataset=None):
        """Testbed::save_snapshot(path, include_optimizer_state=false) (src/testbed.cu:3280-3314): the inference (EMA) parameters as
        binary16, the occupancy grid as binary16, the batch-size controller and the training step, inside `network_config`."""
        from . import snapshot as snap
        ts, rays, _, measured_before = self.get_train_state()
        grid, _ = self.export_density_grid()
        if ts == 0:
            grid = np.zeros(0, np.float32)              # never populated: the reference's grid is still empty
        cfg = snap.build_snapshot(network_config, self.export_params_fp16(use_ema=True), grid, ts, self.last_loss if loss is None else loss, rays, self.last_measured_batch_size, measured_before,
                                  aabb_scale=aabb_scale, dataset=dataset)
        snap.write_snapshot(path, cfg)
        return cfg

    def load_snapshot(self, path):
        """Testbed::load_snapshot (src/testbed.cu:3333-3390): parameters (training, inference and fp32 master all take the stored
        binary16 values, Adam restarts: trainer.h:263-275), occupancy grid + bitfield, controller state, training step."""
        from . import snapshot as snap
        cfg = snap.read_snapshot(path)
        d = snap.parse_snapshot(cfg)
        if d["params_fp16"].size != self.n_params:
            raise RnbError("Can't set params because CPU buffer has the wrong size.")
        self.import_params_fp16(d["params_fp16"])
        if d["density_grid"].size:
            self.import_density_grid(d["density_grid"], 0)
        self.set_train_state(d["training_step"], d["rays_per_batch"], 0, d["measured_batch_size_before_compaction"])
        # neither the canonical step nor the image count of the last occupancy refresh is in the file: the reference resumes with canonical step 0 and
        # (in the process that loads the snapshot before it ever trained) with n_images_for_training_prev 0, so its first step rebuilds the grid from empty
        self.set_canonical_state(0, self._n_images_prev)
        self.last_measured_batch_size = d["measured_batch_size"]
        return cfg

    # --- mesh extraction and export ---
    def _mesh_info(self, info):
        return dict(n_verts=int(info.n_verts), n_verts_padded=int(info.n_verts_padded), n_indices=int(info.n_indices), res=tuple(int(x) for x in info.res),
                    stage_ms=dict(zip(("sdf_sweep", "bits_count_scan", "vertices_normals_faces", "colors"), (round(float(x), 4) for x in info.stage_ms))))

    def marching_cubes(self, res, aabb_min=(0.0, 0.0, 0.0), aabb_max=(1.0, 1.0, 1.0), thresh=0.0, use_ema=True, stream=None):
        """Testbed::marching_cubes (src/testbed_nerf.cu:4297-4348): the mesh stays on the device; returns its sizes."""
        if np.isscalar(res):
            res = (res, res, res)
        r = (C.c_uint32 * 3)(*[int(x) for x in res]); a = (C.c_float * 3)(*[float(x) for x in aabb_min]); b = (C.c_float * 3)(*[float(x) for x in aabb_max])
        info = MeshInfo()
        self._chk(self.L.rnb_marching_cubes(self.h, r, a, b, C.c_float(thresh), int(use_ema), C.c_void_p(stream), C.byref(info)))
        return self._mesh_info(info)

    def marching_cubes_from_density(self, density_dev_ptr, res, aabb_min=(0.0, 0.0, 0.0), aabb_max=(1.0, 1.0, 1.0), thresh=0.0, with_colors=True, use_ema=True, stream=None):
        """marching_cubes_gpu (src/marching_cubes.cu:794-822) + normals (+ colours) on a device lattice density[x + y rx + z rx ry]."""
        r = (C.c_uint32 * 3)(*[int(x) for x in res]); a = (C.c_float * 3)(*[float(x) for x in aabb_min]); b = (C.c_float * 3)(*[float(x) for x in aabb_max])
        info = MeshInfo()
        self._chk(self.L.rnb_marching_cubes_from_density(self.h, C.c_void_p(density_dev_ptr), r, a, b, C.c_float(thresh), int(with_colors), int(use_ema), C.c_void_p(stream), C.byref(info)))
        return self._mesh_info(info)

    def mesh_buffers(self):
        v = C.c_void_p(); n = C.c_void_p(); c = C.c_void_p(); i = C.c_void_p(); info = MeshInfo()
        self._chk(self.L.rnb_mesh_buffers(self.h, C.byref(v), C.byref(n), C.byref(c), C.byref(i), C.byref(info)))
        return dict(verts=v.value, normals=n.value, colors=c.value, indices=i.value, **self._mesh_info(info))

    def mesh_download(self):
        """Host copies: V [n_verts_padded, 3], N (unnormalised area-weighted sums), C, F [n_tris, 3]."""
        m = self.mesh_buffers()
        nv, ni = m["n_verts_padded"], m["n_indices"]
        V = np.zeros((nv, 3), np.float32); N = np.zeros((nv, 3), np.float32); Cc = np.zeros((nv, 3), np.float32); F = np.zeros(ni, np.uint32)
        self._chk(self.L.rnb_mesh_download(self.h, _p(V, C.c_float), _p(N, C.c_float), _p(Cc, C.c_float), _p(F, C.c_uint32)))
        return dict(V=V, N=N, C=Cc, F=F.reshape(-1, 3), n_verts=m["n_verts"])

    def compute_marching_cubes_mesh(self, res, aabb_min=(0.0, 0.0, 0.0), aabb_max=(1.0, 1.0, 1.0), thresh=0.0, use_ema=True):
        """Testbed::compute_marching_cubes_mesh (src/python_api.cu:99-122): dict V, N (normalised), C, F."""
        self.marching_cubes(res, aabb_min, aabb_max, thresh, use_ema)
        m = self.mesh_download()
        z = (m["N"] * m["N"]).sum(1, keepdims=True)
        m["N"] = np.where(z > 0, m["N"] / np.sqrt(np.maximum(z, 1e-45)), m["N"]).astype(np.float32)
        return m

    def save_mesh(self, path, nerf_scale=1.0, nerf_offset=(0.0, 0.0, 0.0), n2w_s=1.0, n2w_t=(0.0, 0.0, 0.0), invert_normals=False, stream=None):
        """save_mesh (src/marching_cubes.cu:824-982) of the current mesh; returns the bytes written."""
        m = self.mesh_buffers()
        return save_mesh_device(path, m["verts"], m["normals"], m["colors"], m["indices"], m["n_verts_padded"], m["n_indices"], nerf_scale, nerf_offset, n2w_s, n2w_t, invert_normals, stream)

    def compute_and_save_marching_cubes_mesh(self, path, res, aabb_min=(0.0, 0.0, 0.0), aabb_max=(1.0, 1.0, 1.0), thresh=0.0, use_ema=True,
                                             nerf_scale=1.0, nerf_offset=(0.0, 0.0, 0.0), n2w_s=1.0, n2w_t=(0.0, 0.0, 0.0), from_na=False):
        """Testbed::compute_and_save_marching_cubes_mesh (src/testbed.cu:369-381); the dataset's scale / offset / n2w come from the caller."""
        info = self.marching_cubes(res, aabb_min, aabb_max, thresh, use_ema)
        self.save_mesh(path, nerf_scale, nerf_offset, n2w_s, n2w_t, invert_normals=from_na)
        return info

    # --- stage-level calls (host buffers) ---
    def stage_generate(self, n_rays, n_rays_total, max_samples):
        ri = np.zeros(n_rays, np.uint32); rays = np.zeros((n_rays, 6), np.float32); ns = np.zeros((n_rays, 2), np.uint32)
        coords = np.zeros((max_samples, 7), np.float32); cnt = (C.c_uint32 * 2)()
        self._chk(self.L.rnb_stage_generate(self.h, n_rays, n_rays_total, max_samples, _p(ri, C.c_uint32), _p(rays, C.c_float), _p(ns, C.c_uint32), _p(coords, C.c_float), cnt))
        k = int(cnt[0])
        return dict(ray_indices=ri[:k], rays=rays[:k], numsteps=ns[:k], coords=coords, n_kept=k, n_samples=int(cnt[1]))

    def stage_forward(self, coords, use_ema=False, want_normal=True):
        coords = np.ascontiguousarray(coords, np.float32); n = coords.shape[0]
        out = np.zeros((n, 16), np.float32); nrm = np.zeros((n, 3), np.float32) if want_normal else None
        self._chk(self.L.rnb_stage_forward(self.h, _p(coords, C.c_float), C.c_size_t(n), int(use_ema), _p(out, C.c_float), _p(nrm, C.c_float)))
        return out, nrm

    def stage_loss(self, out_c, ray_indices, n_fwd, cbase, n_emit, n_rays, n_rays_total):
        out_c = np.ascontiguousarray(out_c, np.float32); k = len(ray_indices)
        dout = np.zeros_like(out_c); lo = np.zeros(k, np.float32); ek = np.zeros(k, np.float32); ml = np.zeros(k, np.float32)
        a = [np.ascontiguousarray(x, np.uint32) for x in (ray_indices, n_fwd, cbase, n_emit)]
        self._chk(self.L.rnb_stage_loss(self.h, _p(out_c, C.c_float), _p(a[0], C.c_uint32), _p(a[1], C.c_uint32), _p(a[2], C.c_uint32), _p(a[3], C.c_uint32), k, n_rays, n_rays_total,
                                        _p(dout, C.c_float), _p(lo, C.c_float), _p(ek, C.c_float), _p(ml, C.c_float)))
        return dout, lo, ek, ml

    def stage_backward(self, coords, dout, n_in_rollover):
        coords = np.ascontiguousarray(coords, np.float32); dout = np.ascontiguousarray(dout, np.float32)
        g = np.zeros(self.n_params, np.float32)
        self._chk(self.L.rnb_stage_backward(self.h, _p(coords, C.c_float), _p(dout, C.c_float), C.c_size_t(coords.shape[0]), n_in_rollover, _p(g, C.c_float)))
        return g

    def stage_optimizer(self, grads=None):
        g = np.ascontiguousarray(grads, np.float32) if grads is not None else None
        self._chk(self.L.rnb_stage_optimizer(self.h, _p(g, C.c_float)))
