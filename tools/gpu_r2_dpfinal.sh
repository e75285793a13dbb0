#!/bin/bash
# final multi-GPU validation of the tree: (N = 2) the 2-rank GPU tests + the default bench; (N = 8) the default bench exactly as the driver launches it
N=${2:-2}; O=gpurun_out/${1:-r2dpfinal}_n$N; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
if [ $N -eq 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_data_parallel.py tests/test_gpu_errors.py -m gpu -q --timeout 800 --timeout-method thread > $O/pytest_dp.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dp.log; tail -4 $O/pytest_dp.log | cut -c1-300
  cp gpurun_out/dp_comm_check_n2.json $O/ 2>/dev/null
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 600 $TR --master-port 29712 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
grep -E "NVLS multicast|nranks $N" $O/bench.err | head -3 | cut -c1-160
python - <<PY
import json
d=json.load(open("$O/bench.json")); print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["clocks"], {k:(round(v["value"]),round(v["ms_per_step"],4)) for k,v in d.get("records",{}).items() if "value" in v}, d["config"]["parallelism"][:120])
PY
