#!/bin/bash
# Runs on the B200 box (via gpurun): GPU parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
# Usage: tools/gpu_check.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-k_backward_mma}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 400 --csv --log-file $O/launches.csv \
    python bench.py --pretrain 300 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 320 -c 2 -o $O/prof_top \
    python bench.py --pretrain 300 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_full_bench.log 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench.json; tail -2 $O/bench.err
