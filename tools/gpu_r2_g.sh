#!/bin/bash
# round 2 session g: training stream with a priority above the library's side stream: A/B of the march start point, then the default bench with records
O=gpurun_out/${1:-r2g}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
run_bench() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-records --steps 150 $BARGS > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"; }
BARGS="--workload normals --pretrain 300"
run_bench n300_hi_at3 RNB_PRELAUNCH_AT=3
run_bench n300_hi_at2 RNB_PRELAUNCH_AT=2
run_bench n300_hi_at1 RNB_PRELAUNCH_AT=1
run_bench n300_def_at3 RNB_BENCH_STREAM=default
BARGS=""
run_bench a700_hi_at3 RNB_PRELAUNCH_AT=3
run_bench a700_hi_at2 RNB_PRELAUNCH_AT=2
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$O/bench_*.json")):
    n=os.path.basename(f)[6:-5]
    try:
        d=json.load(open(f)); print("%-18s"%n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "L", d["config"]["live_hash_levels"], {k[:6]:v["ms"] for k,v in d["roofline"]["stages"].items()})
    except Exception as e: print(n, "failed", e)
PY
