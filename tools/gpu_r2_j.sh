#!/bin/bash
# round 2 session j: canary, optimizer parity, full-length drop-in comparison on the reference-sized dataset (96 views 1600x1200, 10000 steps, mesh 512), bench
O=gpurun_out/${1:-r2j}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py tests/test_gpu_controller.py -m gpu -q --timeout 240 --timeout-method thread > $O/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_subset.log; tail -3 $O/pytest_gpu_subset.log | cut -c1-300
timeout 1500 python tools/dropin_run.py $O/dropin_full --iters 10000 --res 512 --views 96 --width 1600 --height 1200 --timeout 900 > $O/dropin_full.log 2>&1; echo "dropin rc=$?" >> $O/dropin_full.log; tail -6 $O/dropin_full.log | cut -c1-600
rm -rf $O/dropin_full/scene_*
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
timeout 300 python bench.py --no-cpu-baseline --no-records --steps 200 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench.json")); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, round(d["roofline"]["frac"],4))
PY
