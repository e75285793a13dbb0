#!/bin/bash
# A/B of the paired 16-byte gradient atomics (RNB_SCATTER_PAIR) + the GPU suite on the new default
O=gpurun_out/${1:-ab}; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log | cut -c1-300
for v in 1 0 1 0; do
  RNB_SCATTER_PAIR=$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_pair$v.json 2> $O/bench_pair$v.err
  python - <<PY
import json
d=json.load(open("$O/bench_pair$v.json")); print("pair=$v value", round(d["value"]), "ms", round(d["ms_per_step"],4), "backward", d["roofline"]["stages"]["backward"]["ms"], "passA", d["roofline"]["stages"]["pass_a_sdf_normal"]["ms"])
PY
done
