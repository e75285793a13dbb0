"""The C-ABI library loads and exports every symbol include/rnb_b200.h declares (no compute, CPU only)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rnb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rnb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    names = _declared()
    assert len(names) >= 30
    L = pkg.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(pkg.EXPORTED_SYMBOLS) == names
    assert L.rnb_abi_version() == 1


def test_struct_sizes_match_header(pkg):
    # the ctypes mirrors must match the C layout: plain 4-byte fields (and two pointers in rnb_view)
    assert C.sizeof(pkg.Config) == 31 * 4
    assert C.sizeof(pkg.Flags) == 12 * 4
    assert C.sizeof(pkg.StepStats) == 11 * 4
    assert C.sizeof(pkg.View) == 2 * 8 + 6 * 4 + 12 * 4


def test_error_path_without_gpu_or_bad_config(pkg):
    L = pkg.lib()
    cfg = pkg.default_config(n_levels=40)
    h = C.c_void_p()
    rc = L.rnb_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and len(L.rnb_last_error()) > 0      # invalid config (or no device): loud failure, never a fallback


def test_header_is_plain_c_and_links_from_c(pkg, tmp_path):
    """the boundary is a C ABI: the header must compile as C99 (and as C++), and a C program using it must link against the library and
    get the loud failure path without a device / with a bad configuration"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc"); gxx = shutil.which("g++")
    if not gcc or not gxx:
        import pytest
        pytest.skip("no host compiler")
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "cabi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "rnb_b200.h"
int main(void) {
	rnb_config cfg; rnb_flags fl; rnb_ctx* ctx = 0; rnb_raymesh* rm = 0;
	rnb_default_config(&cfg); rnb_default_flags(&fl);
	if (rnb_abi_version() != RNB_ABI_VERSION) return 2;
	cfg.n_levels = 40;                                   /* invalid: more levels than the library supports */
	if (rnb_create(&cfg, &ctx) == RNB_OK) return 3;
	if (strlen(rnb_last_error()) == 0) return 4;
	if (rnb_train(0, 0, 0) == RNB_OK) return 5;           /* null context */
	if (rnb_raymesh_create(0, 0, 0, 0, 0, &rm) == RNB_OK) return 6;
	if (rnb_raymesh_destroy(0) != RNB_OK) return 7;
	printf("abi %u ok\n", rnb_abi_version());
	return 0;
}
''')
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    exe = tmp_path / "cabi"
    libdir = os.path.join(ROOT, "rnb-neus2_b200")
    subprocess.check_call([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lrnb_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi 1 ok" in out.stdout


def test_create_rejects_configurations_that_would_misbehave_on_the_device(pkg):
    """shifts past the word size, zero divisors, NaN resolutions: stopped in rnb_create with RNB_ERR_INVALID and a message, before any
    device call (valid configurations get past the checks and fail here only for the missing device)"""
    import torch
    L = pkg.lib()
    bad = [dict(log2_hashmap_size=40), dict(log2_hashmap_size=0), dict(base_resolution=0), dict(target_batch_size=0), dict(target_batch_size=1 << 31),
           dict(rays_per_batch=0), dict(rays_per_batch=1 << 19), dict(top_resolution=-5.0), dict(top_resolution=float("nan")), dict(per_level_scale=float("nan")),
           dict(per_level_scale=64.0), dict(loss_scale=0.0), dict(beta1=1.0), dict(beta2=-0.1), dict(epsilon=0.0), dict(ema_decay=1.0), dict(learning_rate=float("inf")),
           dict(density_grid_decay=0.0), dict(sdf_bias=float("nan")), dict(world_size=4, target_batch_size=2)]
    for kw in bad:
        h = C.c_void_p()
        rc = L.rnb_create(C.byref(pkg.default_config(**kw)), C.byref(h))
        assert rc == -1 and len(L.rnb_last_error()) > 0, kw
    if not torch.cuda.is_available():
        good = [dict(), dict(n_levels=1), dict(per_level_scale=1.3819), dict(n_levels=8, log2_hashmap_size=14, sdf_n_neurons=32, rgb_n_neurons=32, rgb_n_hidden_layers=1, rays_per_batch=128),
                dict(world_size=8, rank=7, target_batch_size=(1 << 18) * 8), dict(lr_decay_interval=0, lr_decay_base=0.0), dict(ema_decay=0.0, l2_reg=0.0, learning_rate=0.0)]
        for kw in good:
            h = C.c_void_p()
            rc = L.rnb_create(C.byref(pkg.default_config(**kw)), C.byref(h))
            assert rc == -2, (kw, L.rnb_last_error())          # RNB_ERR_CUDA: no device in this container, and no fallback


def test_communicator_entry_points_reject_misuse(pkg):
    """rnb_comm_*: NCCL is opened at run time; null contexts / ids are refused with a message and nothing crashes without a device"""
    L = pkg.lib()
    ident = (C.c_uint8 * 128)()
    assert L.rnb_comm_init(None, ident) == -1 and b"null" in L.rnb_last_error()
    assert L.rnb_comm_adopt(None, None) == -1
    assert L.rnb_comm_destroy(None) == -1
    assert L.rnb_comm_info(None, None) == -1
    assert L.rnb_comm_unique_id(None) == -1
    assert L.rnb_set_canonical_state(None, 0, 0) == -1
