#!/bin/bash
# round 2 session k: A/B of the cached Adam learning rate (same box, alternating)
O=gpurun_out/${1:-r2k}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
run_bench() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-records --steps 200 $BARGS > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"; }
BARGS=""
run_bench lrc1_a RNB_ADAM_LRCACHE=1
run_bench lrc0_a RNB_ADAM_LRCACHE=0
run_bench lrc1_b RNB_ADAM_LRCACHE=1
run_bench lrc0_b RNB_ADAM_LRCACHE=0
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$O/bench_*.json")):
    n=os.path.basename(f)[6:-5]
    try:
        d=json.load(open(f)); print("%-10s"%n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k[:6]:v["ms"] for k,v in d["roofline"]["stages"].items()})
    except Exception as e: print(n, "failed", e)
PY
