"""Static evidence in the built library (cuobjdump -sass, tools/sass_evidence.py): the default-path kernels use the Blackwell tensor path
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st) and not the legacy one (HMMA), the backward carries the vector reductions and the
warp aggregation, the cross-check kernels are the mma.sync generation.  profiles/r02_sass_mnemonics.txt is the committed copy."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_default_kernels_use_tcgen05_and_vector_reductions(capsys):
    import sass_evidence
    counts, pretty = sass_evidence.main()
    capsys.readouterr()
    by_name = {pretty[k]: v for k, v in counts.items()}
    for name in ("tc::k_sdf_tc<64, 0, false>", "tc::k_full_tc<64, true>", "tc::k_backward_tc<64, true, 1>", "tc::k_sdf_tc<64, 2, false>", "tc::k_backward_tc<32, false, 1>", "tc::k_backward_tc<64, true, 2>"):
        c = by_name[name]
        assert c["UTCHMMA"] > 0 and c["LDTM"] > 0 and c["STTM"] > 0 and c["UTCBAR"] > 0, (name, dict(c))
        assert c["HMMA"] == 0, name
    bw = by_name["tc::k_backward_tc<64, true, 1>"]
    assert bw["REDG.x4"] > 0 and bw["REDG.x2"] > 0          # 16-byte pair reductions and the 8-byte ones
    assert bw["REDUX"] > 0 and bw["SHFL"] >= 17             # redux.max of the run length, shfl.up of the key + 16 shfl.down per step
    mma = by_name["k_backward_mma<64, 64, true>"]
    assert mma["HMMA"] > 0 and mma["UTCHMMA"] == 0 and mma["UBLKCP/UTMA"] > 0          # previous generation: mma.sync + bulk-async weight load
    assert by_name["k_march"]["REDG"] == 0                  # ordered scans instead of arrival-order atomics
    assert bw["USETMAXREG"] >= 2                            # warp-specialised: chain warps grow, scatter warps shrink their register budget (setmaxnreg)
    committed = open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt")).read()
    assert "tc::k_backward_tc<64, true, 1>" in committed and "UTCHMMA" in committed and "USETMAXREG" in committed
