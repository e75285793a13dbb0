#!/bin/bash
# round 2 session e: canary, whole GPU suite (no -x), A/B of the asynchronous step end / march pre-launch, default bench with both scatter configurations
O=gpurun_out/${1:-r2e}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -12 $O/pytest_gpu.log | cut -c1-300
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
run_bench() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-records --steps 150 $BARGS > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"; }
BARGS="--workload normals --pretrain 300"
run_bench n300_default RNB_X=0
run_bench n300_sync RNB_ASYNC_END=0
run_bench n300_nopre RNB_PRELAUNCH=0
run_bench n300_nopre_sync RNB_PRELAUNCH=0 RNB_ASYNC_END=0
run_bench n300_pre0 RNB_PRELAUNCH_AT=0
run_bench n300_wg2 RNB_BW_SCATTER_WG=2
BARGS=""
run_bench a700_wg1 RNB_BW_SCATTER_WG=1
run_bench a700_wg2 RNB_BW_SCATTER_WG=2
run_bench a700_stageA1 RNB_STAGE_LEVELS_A=1
run_bench a700_nostage RNB_STAGE_LEVELS=0
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$O/bench_*.json")):
    n=os.path.basename(f)[6:-5]
    try:
        d=json.load(open(f)); print("%-18s"%n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "L", d["config"]["live_hash_levels"], {k[:6]:v["ms"] for k,v in d["roofline"]["stages"].items()})
    except Exception as e: print(n, "failed", e)
PY
