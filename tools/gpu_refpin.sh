#!/bin/bash
# On the B200 box: pin oracle + CUDA path against the reference build (oracle/_ref); optional reference timing.
O=gpurun_out/refpin
mkdir -p $O
if [ -x tests/cuda/umma_probe ]; then timeout 120 tests/cuda/umma_probe > $O/umma_probe.log 2>&1; echo "rc=$?" >> $O/umma_probe.log; cat $O/umma_probe.log; fi
timeout 900 python tests/ref_pin.py --config small --out $O > $O/pin_small.log 2>&1; echo "rc=$?" >> $O/pin_small.log
timeout 900 python tests/ref_pin.py --config small --albedo --out $O > $O/pin_small_alb.log 2>&1; echo "rc=$?" >> $O/pin_small_alb.log
timeout 1500 python tests/ref_pin.py --config full --steps 34 --out $O > $O/pin_full.log 2>&1; echo "rc=$?" >> $O/pin_full.log
if [ "$1" == "time" ]; then timeout 900 python tools/ref_time.py --steps 600 --pin-rays 4096 --out $O/ref_time.json > $O/ref_time.log 2>&1; echo "rc=$?" >> $O/ref_time.log; fi
tail -2 $O/pin_small.log | cut -c1-600; tail -2 $O/pin_small_alb.log | cut -c1-300; tail -2 $O/pin_full.log | cut -c1-600
