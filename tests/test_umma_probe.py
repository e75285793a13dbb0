"""Hardware check of the tcgen05 operand layouts the network kernels rely on (tests/cuda/umma_probe.cu, built by
__graft_entry__.build()): K-major / MN-major shared-memory descriptors over the panel layout, the A operand in TMEM, and
the TMEM lane map of an M = 64 accumulator."""
import os
import subprocess
import pytest

BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda", "umma_probe")


@pytest.mark.gpu
def test_umma_layouts_on_hardware():
    if not os.path.exists(BIN):
        pytest.skip("tests/cuda/umma_probe not built (run __graft_entry__.build())")
    out = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120).stdout
    assert "UMMA PROBE: ALL OK" in out, out
    assert "r16->32" in out and "r48->96" in out        # M = 64: row r lives in lane (r / 16) * 32 + r % 16
