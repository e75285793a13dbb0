#!/bin/bash
# Round profile set: launch list (all kernels of ~25 steps) + ncu --set full of every hot kernel of ONE step, + clocks during a plain bench.
TAG=${1:-r01_final}
O=gpurun_out/$TAG; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 400 --csv --log-file $O/launches.csv \
    python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_full_tc|k_backward_tc|k_adam_ema|k_march|k_loss|k_sdf_tc|k_compact_count" -s 4480 -c 8 -o $O/prof_step \
    python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_step.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err
kill $SMI
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
ls -la $O; cat $O/bench.json | cut -c1-600; cat $O/bench_reference.json | cut -c1-600
