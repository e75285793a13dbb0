#!/bin/bash
# A/B of the warp-aggregated scatter (RNB_SCATTER_AGG=n levels) and of an early read-back of the step counters (RNB_ASYNC_END; it was
# slower and has been removed from the library since: profiles/r01_ab_scatter_agg.txt) + the GPU suite under both
# settings + the new albedo-stage tests.  One call, ~6 min.
O=gpurun_out/${1:-ab2}; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_albedo.py -q -x -s > $O/pytest_albedo.log 2>&1; echo "albedo rc=$?" >> $O/pytest_albedo.log; grep -a "raymesh full size\|passed\|failed\|rc=\|Error\|assert" $O/pytest_albedo.log | cut -c1-300 | tail -12
timeout 300 python -m pytest tests -m gpu -q -k "not albedo and not raymesh" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log | cut -c1-300
RNB_SCATTER_AGG=8 RNB_ASYNC_END=1 timeout 300 python -m pytest tests -m gpu -q -k "not albedo and not raymesh" > $O/pytest_gpu_flags.log 2>&1; echo "pytest(flags) rc=$?" >> $O/pytest_gpu_flags.log; tail -4 $O/pytest_gpu_flags.log | cut -c1-300
for cfg in "0 0" "8 0" "5 0" "0 1" "8 1" "6 1" "0 0" "8 1"; do
  set -- $cfg
  RNB_SCATTER_AGG=$1 RNB_ASYNC_END=$2 timeout 150 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_agg$1_async$2.json 2> $O/bench_agg$1_async$2.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_agg$1_async$2.json")); st=d["roofline"]["stages"]
    print("agg=$1 async=$2 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "backward", st["backward"]["ms"], "passA", st["pass_a_sdf_normal"]["ms"], "adam", st["adam_ema"]["ms"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("agg=$1 async=$2 FAILED", e)
PY
done
