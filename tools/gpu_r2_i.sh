#!/bin/bash
# round 2 session i: canary, GPU suite (default and with the pipelined gather), A/B of the pipelined gather, FULL + albedo reference pin
O=gpurun_out/${1:-r2i}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log | cut -c1-300
RNB_GATHER_PIPE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_controller.py tests/test_reference_golden.py -m gpu -q --timeout 240 --timeout-method thread > $O/pytest_gpu_pipe.log 2>&1; echo "pytest pipe rc=$?" >> $O/pytest_gpu_pipe.log; tail -4 $O/pytest_gpu_pipe.log | cut -c1-300
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
run_bench() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-records --steps 200 $BARGS > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"; }
BARGS=""
run_bench a700_pipe0 RNB_GATHER_PIPE=0
run_bench a700_pipe1 RNB_GATHER_PIPE=1
run_bench a700_pipe0b RNB_GATHER_PIPE=0
run_bench a700_pipe1b RNB_GATHER_PIPE=1
BARGS="--workload normals --pretrain 300"
run_bench n300_pipe0 RNB_GATHER_PIPE=0
run_bench n300_pipe1 RNB_GATHER_PIPE=1
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$O/bench_*.json")):
    n=os.path.basename(f)[6:-5]
    try:
        d=json.load(open(f)); print("%-14s"%n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "L", d["config"]["live_hash_levels"], {k[:6]:v["ms"] for k,v in d["roofline"]["stages"].items()}, round(d["roofline"]["frac"],4))
    except Exception as e: print(n, "failed", e)
PY
timeout 700 python tests/ref_pin.py --config full --albedo --steps 700 --dump-steps 0,1,300,699 --out $O/refpin_full_alb > $O/refpin_full_alb.log 2>&1; echo "refpin rc=$?" >> $O/refpin_full_alb.log; tail -3 $O/refpin_full_alb.log | cut -c1-700
rm -f $O/refpin_full_alb/golden_full_probe.npz
