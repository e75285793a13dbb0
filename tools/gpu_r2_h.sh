#!/bin/bash
# round 2 session h: canary, GPU suite, run_pipeline.py --has-albedo against testbed_rnb (BASELINE configs[2]), both bench arms of the tree
O=gpurun_out/${1:-r2h}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 900 python tools/dropin_run.py $O/dropin_albedo --iters 3000 --res 256 --only rnb --skip-two-stage --pipeline-albedo > $O/dropin_albedo.log 2>&1; echo "dropin albedo rc=$?" >> $O/dropin_albedo.log; tail -4 $O/dropin_albedo.log | cut -c1-500
rm -rf $O/dropin_albedo/rnb_input_albedo $O/dropin_albedo/pipeline_albedo_rnb/prepared_data $O/dropin_albedo/pylib
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/bench_reference.json
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench.json")); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}); print(d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["roofline"]["step"], d["clocks"], d.get("cpu_baseline",{}).get("value"), d["gpu_launches"])
print(json.dumps(d.get("records"))[:1500])
PY
