#!/bin/bash
# A/B of the point behind which the next step's march may start (RNB_PRELAUNCH_AT: 0 backward, 1 loss, 2 pass A, 3 scan/emit)
O=gpurun_out/${1:-ab3}; mkdir -p $O
RNB_PRELAUNCH_AT=3 timeout 300 python -m pytest tests -m gpu -q -k "not albedo and not raymesh" > $O/pytest_gpu_at3.log 2>&1; echo "pytest(at3) rc=$?" >> $O/pytest_gpu_at3.log; tail -4 $O/pytest_gpu_at3.log | cut -c1-300
for at in 0 1 2 3 0 1 2 3; do
  RNB_PRELAUNCH_AT=$at timeout 150 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_at${at}.json 2> $O/bench_at${at}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_at${at}.json")); st=d["roofline"]["stages"]
    print("at=$at value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "backward", st["backward"]["ms"], "passA", st["pass_a_sdf_normal"]["ms"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("at=$at FAILED", e)
PY
done
