#!/bin/bash
# N-GPU session (gpurun --gpus N): canary, data-parallel check of the library-owned communicator, bench at N GPUs (all-reduce and sharded)
N=${2:-2}; O=gpurun_out/${1:-r2dp}_n$N; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${3:-check}" == "check" ]; then
  RNB_CHECK_STEPS=30 timeout 600 $TR --master-port 29611 tools/dp_comm_check.py > $O/dp_comm_check.log 2>&1; echo "dp_comm_check rc=$?"; grep -E "^\{" $O/dp_comm_check.log | tail -1 | cut -c1-1500
  cp gpurun_out/dp_comm_check_n$N.json $O/ 2>/dev/null
fi
NCCL_DEBUG=INFO timeout 500 $TR --master-port 29612 bench.py --gpus $N --no-cpu-baseline --steps 200 > $O/bench_allreduce.json 2> $O/bench_allreduce.err; echo "bench allreduce rc=$?"
grep -E "NVLS|Connected all|nranks|ncclCommInitRank" $O/bench_allreduce.err | head -8 | cut -c1-200
RNB_DP=sharded timeout 400 $TR --master-port 29613 bench.py --gpus $N --no-cpu-baseline --no-records --steps 200 > $O/bench_sharded.json 2> $O/bench_sharded.err; echo "bench sharded rc=$?"
python - <<PY
import json
for n in ("bench_allreduce","bench_sharded"):
    try:
        d=json.load(open("$O/%s.json"%n)); print(n, "N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["clocks"], {k:(round(v["value"]),round(v["ms_per_step"],4)) for k,v in d.get("records",{}).items() if "value" in v})
    except Exception as e: print(n, "failed", e)
PY
