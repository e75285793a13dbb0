"""Error behaviour of the C ABI on the GPU box: every misuse is a negative status with a message (the caller turns it into the
reference's std::runtime_error), never a crash and never a silent fallback."""
import ctypes as C
import numpy as np
import pytest
from common import SMALL, product_config

pytestmark = pytest.mark.gpu


def _err(pkg, rc):
    assert rc < 0
    return pkg.lib().rnb_last_error().decode()


def test_misuse_returns_status_and_message(pkg, tmp_path):
    import torch
    L = pkg.lib()
    t = pkg.Testbed(product_config(pkg, SMALL)); t.init_params()
    st = pkg.StepStats()
    assert "dataset" in _err(pkg, L.rnb_train_step(t.h, None, C.byref(st)))                  # training without a dataset
    assert _err(pkg, L.rnb_train_step(None, None, C.byref(st)))
    p = np.zeros(t.n_params - 1, np.float32)
    assert "parameter" in _err(pkg, L.rnb_set_params_fp32(t.h, p.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(p.size)))
    g = np.zeros(5, np.float32)
    assert "128^3" in _err(pkg, L.rnb_import_density_grid(t.h, g.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(g.size), 0))
    assert _err(pkg, L.rnb_upload_dataset(t.h, None, 0))
    # mesh path
    d = torch.zeros(24 * 8 * 8, device="cuda")
    res = (C.c_uint32 * 3)(24, 8, 8); mn = (C.c_float * 3)(0, 0, 0); mx = (C.c_float * 3)(1, 1, 1); info = pkg.MeshInfo()
    assert "multiple of 16" in _err(pkg, L.rnb_marching_cubes_from_density(t.h, C.c_void_p(d.data_ptr()), res, mn, mx, C.c_float(0), 0, 0, None, C.byref(info)))
    assert "no mesh" in _err(pkg, L.rnb_mesh_download(t.h, None, None, None, None))
    big = (C.c_uint32 * 3)(4096, 4096, 4096)
    assert "2^32" in _err(pkg, L.rnb_marching_cubes(t.h, big, mn, mx, C.c_float(0), 0, None, C.byref(info)))
    v = torch.zeros(128 * 3, device="cuda"); i = torch.zeros(4, dtype=torch.int32, device="cuda")
    nb = C.c_uint64()
    args = (C.c_float(1.0), (C.c_float * 3)(0, 0, 0), C.c_float(1.0), (C.c_float * 3)(0, 0, 0), 0, None, C.byref(nb))
    assert "multiple of 3" in _err(pkg, L.rnb_save_mesh(C.c_void_p(v.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(i.data_ptr()), 128, 4, str(tmp_path / "m.obj").encode(), *args))
    assert "Failed to open" in _err(pkg, L.rnb_save_mesh(C.c_void_p(v.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(i.data_ptr()), 128, 3, str(tmp_path / "no_such_dir" / "m.obj").encode(), *args))
    # the context is still usable afterwards
    info = t.marching_cubes_from_density(torch.full((16 * 16 * 16,), 1.0, device="cuda").data_ptr(), (16, 16, 16), with_colors=False)
    assert info["n_verts"] == 0 and info["n_indices"] == 0
    t.close()


def test_data_parallel_context_without_communicator_fails_loudly(pkg, scene_mod):
    """world_size > 1: rnb_train refuses to run a step whose gradients nobody would exchange (no silent single-rank training); the split entry points
    (the caller's own collective) keep working, and the stale-EMA guard of the sharded optimizer is not armed without a communicator"""
    views = scene_mod.make_scene(4, 64, 64, with_albedo=False)
    t = pkg.Testbed(pkg.default_config(n_levels=8, log2_hashmap_size=14, sdf_n_neurons=32, rgb_n_neurons=32, rgb_n_hidden_layers=1, rays_per_batch=256, world_size=2, rank=0))
    t.init_params(); t.load_training_data(views)
    with pytest.raises(pkg.RnbError, match="rnb_comm_init"):
        t.train()
    assert t.comm_info() == dict(installed=False, nccl_version=t.comm_info()["nccl_version"], sharded=False, one_sample_order=False, world_size=2)
    t.training_prep_nerf()
    t.train_step_begin()
    st = t.train_step_end()                      # this rank's half of the rays, unreduced: the caller would have all-reduced rnb_grad_buffer in between
    assert st.training_step == 1 and st.n_rays == 256 and st.n_rays_kept <= 128
    t.export_params_fp16(use_ema=True)           # no communicator, no sharded optimizer: the inference parameters are readable
