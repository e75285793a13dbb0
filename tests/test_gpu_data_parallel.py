"""Data parallelism BEHIND the C ABI on hardware (needs >= 2 GPUs on the box; skipped otherwise): tools/dp_comm_check.py under torchrun —
rnb_comm_init + rnb_train with the library's own NCCL communicator (binary16 all-reduce and the sharded optimizer) against a single-GPU run of
the same global batch and against the external fp32 all-reduce protocol; replicas bit-identical, gradient buffer clean, adaptive controller
consistent across ranks."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_communicator_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    env = dict(os.environ); env["RNB_CHECK_STEPS"] = "24"
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tools", "dp_comm_check.py")], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, (p.stdout[-1500:], p.stderr[-1500:])
    out = json.loads(lines[-1])
    assert out["ok"] and all(out["ranks_identical"].values()) and out["adaptive_rays_same_on_all_ranks"]
