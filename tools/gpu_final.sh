#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke, both bench arms
O=gpurun_out/${1:-final}; mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-250 $O/bench_reference.json
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python tools/bench_cmp.py $O 2>/dev/null | head -0; python - <<PY
import json
d=json.load(open("$O/bench.json")); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}); print(d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"])
PY
