#!/bin/bash
# N-GPU session (gpurun --gpus N): one sample order over all ranks (prefix all-gathers) — data-parallel check, then bench A/B against the per-rank rule
N=${2:-2}; O=gpurun_out/${1:-r2exact}_n$N; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${3:-check}" == "check" ]; then
  RNB_CHECK_STEPS=30 timeout 600 $TR --master-port 29611 tools/dp_comm_check.py > $O/dp_comm_check.log 2>&1; echo "dp_comm_check rc=$?"; grep -E "^\{" $O/dp_comm_check.log | tail -1 | cut -c1-2500
  cp gpurun_out/dp_comm_check_n$N.json $O/ 2>/dev/null
  tail -5 $O/dp_comm_check.log | cut -c1-300
fi
RNB_DP_EXACT=1 timeout 400 $TR --master-port 29612 bench.py --gpus $N --no-cpu-baseline --no-records --steps ${4:-200} > $O/bench_one_order.json 2> $O/bench_one_order.err; echo "bench one-order rc=$?"
RNB_DP_EXACT=0 timeout 400 $TR --master-port 29613 bench.py --gpus $N --no-cpu-baseline --no-records --steps ${4:-200} > $O/bench_per_rank.json 2> $O/bench_per_rank.err; echo "bench per-rank rc=$?"
python - <<PY
import json
for n in ("bench_one_order","bench_per_rank"):
    try:
        d=json.load(open("$O/%s.json"%n)); print(n, "N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), d["config"].get("nccl"), {k:v["ms"] for k,v in d["roofline"]["stages"].items()} if "stages" in d["roofline"] else "")
    except Exception as e: print(n, "failed", e)
PY
