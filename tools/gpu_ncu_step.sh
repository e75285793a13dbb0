#!/bin/bash
# one ncu --set full capture of every hot kernel of ONE training step (a single bench run)
TAG=${1:-r01_step}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_full_tc|k_backward_mma|k_adam_ema|k_march|k_loss|k_sdf_tc|k_compact_count" -s 4480 -c 8 -o $O/prof_step \
    python bench.py --pretrain 600 --warmup 3 --steps 30 --no-cpu-baseline > $O/ncu_step.log 2>&1
tail -3 $O/ncu_step.log | cut -c1-200; ls -la $O
