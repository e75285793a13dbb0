#!/bin/bash
# Round-end validation on the final tree, in one call: GPU tests, smoke, both bench arms, then the profile set of the same bench command
# (ncu launch list; ncu --set full of one backward and one pass-A launch).
O=gpurun_out/${1:-final2}; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
( nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 200 > $O/clocks.csv 2>/dev/null & echo $! > $O/smi.pid )
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
kill $(cat $O/smi.pid) 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/bench.json")); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()}); print(d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["clocks"], d["cpu_baseline"]["value"], d["gpu_launches"], d["config"]["network_path"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 400 --csv --log-file $O/launches.csv \
    python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l $O/launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_backward_tc|k_sdf_tc" -s 1240 -c 2 -o $O/prof_bwd_passa \
    python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la $O | head -20
