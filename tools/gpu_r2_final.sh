#!/bin/bash
# round-2 final validation of the tree in one call: GPU suite, smoke, both bench arms, ncu launch list of the bench command
O=gpurun_out/${1:-r2final}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 400 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
( nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 200 > $O/clocks.csv 2>/dev/null & echo $! > $O/smi.pid )
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
kill $(cat $O/smi.pid) 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/bench.json")); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}); print(d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["roofline"]["step"], d["clocks"], d["cpu_baseline"]["value"], d["gpu_launches"])
print({k:(round(v["value"]),round(v["ms_per_step"],4)) for k,v in d.get("records",{}).items() if "value" in v}, d["records"].get("mesh_1024",{}).get("wall_s"))
PY
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 10600 -c 400 --csv --log-file $O/launches.csv python bench.py --pretrain 700 --warmup 3 --steps 30 --no-cpu-baseline --no-records > $O/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l $O/launches.csv
python tools/ncu_launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; head -12 $O/launches_summary.txt
