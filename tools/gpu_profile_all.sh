#!/bin/bash
# Round profile set: launch list (all kernels of ~25 steps) + one ncu --set full capture of each hot kernel.
TAG=${1:-r01_final}
O=gpurun_out/$TAG; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 400 --csv --log-file $O/launches.csv \
    python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1
for K in k_sdf_tc k_full_tc k_backward_mma k_adam_ema k_march k_loss; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 700 -c 1 -o $O/prof_$K \
      python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_$K.log 2>&1
done
ls -la $O
