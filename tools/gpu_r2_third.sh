#!/bin/bash
# round 2, third GPU session: canary, GPU suite on the warp-specialised backward + async step + resume semantics; resume probe; FULL reference pin with
# 14 live levels; bench A/B of the scatter warpgroups and of the staged level in pass A
O=gpurun_out/${1:-r2c}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
RNB_BW_SCATTER_WG=2 timeout 150 python tools/sanitize_case.py network > $O/canary_wg2.log 2>&1 || { echo "CANARY wg2 FAILED"; tail -3 $O/canary_wg2.log; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 240 --timeout-method thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -30 $O/pytest_gpu.log | cut -c1-400
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 400 python tools/resume_probe.py $O/resume > $O/resume.log 2>&1; echo "resume rc=$?" >> $O/resume.log; tail -10 $O/resume.log | cut -c1-700
timeout 600 python tests/ref_pin.py --config full --steps 700 --dump-steps 0,1,300,660,699 --out $O/refpin_full14 > $O/refpin_full14.log 2>&1; echo "refpin rc=$?" >> $O/refpin_full14.log; tail -3 $O/refpin_full14.log | cut -c1-900
run_bench() { # tag, env..., args
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-records --steps 200 $BARGS > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "bench $tag rc=$?"
}
BARGS="--workload normals --pretrain 300" run_bench wg1_normals_p300 RNB_BW_SCATTER_WG=1
BARGS="--workload normals --pretrain 300" run_bench wg2_normals_p300 RNB_BW_SCATTER_WG=2
BARGS="" run_bench wg1 RNB_BW_SCATTER_WG=1
BARGS="" run_bench wg2 RNB_BW_SCATTER_WG=2
BARGS="" run_bench wg1_stageA1 RNB_BW_SCATTER_WG=1 RNB_STAGE_LEVELS_A=1
python - <<PY
import json
for n in ("wg1_normals_p300","wg2_normals_p300","wg1","wg2","wg1_stageA1"):
    try:
        d=json.load(open("$O/bench_%s.json"%n)); print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "levels", d["config"]["live_hash_levels"], {k:v["ms"] for k,v in d["roofline"]["stages"].items()}, d["clocks"])
    except Exception as e: print(n, "failed", e)
PY
