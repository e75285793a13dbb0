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
