#!/bin/bash
# round 2 profile session: ncu launch list of the bench command, ncu --set full of one pass A + one backward launch (+ Adam), compute-sanitizer passes
O=gpurun_out/${1:-r2prof}; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
BENCH="python bench.py --pretrain 700 --warmup 3 --steps 30 --no-cpu-baseline --no-records"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 11300 -c 400 --csv --log-file $O/launches.csv $BENCH > $O/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l $O/launches.csv
python tools/ncu_launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; head -14 $O/launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_backward_tc|k_sdf_tc|k_adam_ema" -s 2190 -c 3 -o $O/prof_full $BENCH > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/prof_full.ncu-rep --page raw --csv > $O/prof_full_raw.csv 2>/dev/null; python tools/ncu_summary.py $O/prof_full_raw.csv > $O/prof_full_summary.txt 2>&1; grep -E "Kernel Name|gpu__time_duration|dram__bytes|issue_active|warps_active|registers_per_thread" $O/prof_full_summary.txt | cut -c1-160
ncu -i $O/prof_full.ncu-rep --page details --csv > $O/prof_full_details.csv 2>/dev/null
bash tools/gpu_sanitize.sh ${1:-r2prof} all
