#!/bin/bash
# ncu full capture of one kernel during a short bench run: tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
TAG=$1; KRE=$2; SKIP=${3:-320}; CNT=${4:-1}
O=gpurun_out/$TAG; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o $O/prof \
    python bench.py --pretrain 300 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -2 $O/ncu_bench.log | cut -c1-300
