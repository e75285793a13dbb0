#!/bin/bash
# round 2, third GPU session: GPU suite on the warp-specialised backward + async step + resume semantics; resume probe; bench A/B of the scatter warpgroups
O=gpurun_out/${1:-r2c}; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -30 $O/pytest_gpu.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python tools/resume_probe.py $O/resume > $O/resume.log 2>&1; echo "resume rc=$?" >> $O/resume.log; tail -10 $O/resume.log | cut -c1-700
for wg in 1 2; do
  RNB_BW_SCATTER_WG=$wg timeout 400 python bench.py --no-cpu-baseline --steps 200 > $O/bench_wg$wg.json 2> $O/bench_wg$wg.err; echo "bench wg$wg rc=$?"
  RNB_BW_SCATTER_WG=$wg timeout 400 python bench.py --no-cpu-baseline --steps 200 --pretrain 700 > $O/bench_wg${wg}_p700.json 2> $O/bench_wg${wg}_p700.err; echo "bench wg$wg p700 rc=$?"
done
python - <<PY
import json
for n in ("bench_wg1","bench_wg1_p700","bench_wg2","bench_wg2_p700"):
    try:
        d=json.load(open("$O/%s.json"%n)); print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "levels", d["config"]["live_hash_levels"], {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["clocks"])
    except Exception as e: print(n, "failed", e)
PY
