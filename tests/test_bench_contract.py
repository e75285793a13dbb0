"""bench.py contract on CPU: the reference arm's fallback (the CPU restatement; used when oracle/_ref or a GPU is absent) prints
exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_cpu_port_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-port", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=580, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training_rays_per_second" and d["unit"] == "rays/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback for the product: without a CUDA device bench.py exits with an error instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_network_path_label_follows_the_kernel_selection(monkeypatch):
    """config.network_path names the kernels the library selects from the same environment variables (csrc/rnb_api.cu)"""
    import importlib
    bench = importlib.import_module("bench")
    for var in ("RNB_NETWORK", "RNB_BACKWARD"):
        monkeypatch.delenv(var, raising=False)
    assert bench._network_path() == "tcgen05 forward + tcgen05 backward"
    monkeypatch.setenv("RNB_BACKWARD", "mma")
    assert bench._network_path() == "tcgen05 forward + mma.sync backward"
    monkeypatch.setenv("RNB_NETWORK", "mma")
    assert "mma.sync" in bench._network_path() and "cross-check" in bench._network_path()
    monkeypatch.setenv("RNB_NETWORK", "simt")
    assert "CUDA-core" in bench._network_path()
    assert bench.DP_MODE_DEFAULT == "sharded"            # the measured default at N = 2 and N = 8 (profiles/r02_dp_n*.txt)
