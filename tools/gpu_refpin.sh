#!/bin/bash
# On the B200 box: pin oracle + CUDA path against the reference build (oracle/_ref), then time the reference CUDA path.
O=gpurun_out/refpin
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python tests/ref_pin.py --config small --steps 6 --out $O > $O/pin_small.log 2>&1; echo "rc=$?" >> $O/pin_small.log
timeout 900 python tests/ref_pin.py --config small --steps 4 --albedo --out $O > $O/pin_small_alb.log 2>&1; echo "rc=$?" >> $O/pin_small_alb.log
timeout 1500 python tests/ref_pin.py --config full --steps 4 --out $O > $O/pin_full.log 2>&1; echo "rc=$?" >> $O/pin_full.log
timeout 900 python tools/ref_time.py --steps 600 --pin-rays 4096 --out $O/ref_time.json > $O/ref_time.log 2>&1; echo "rc=$?" >> $O/ref_time.log
tail -3 $O/pytest_gpu.log; tail -2 $O/pin_small.log | cut -c1-600; tail -2 $O/pin_small_alb.log | cut -c1-300; tail -2 $O/pin_full.log | cut -c1-600; tail -3 $O/ref_time.log | cut -c1-1500
