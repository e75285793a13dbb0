#!/bin/bash
# quick GPU check: parity tests, A/B bench of the network paths
O=gpurun_out/${1:-quick}
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_tc.json 2> $O/bench_tc.err; echo "rc=$?" >> $O/bench_tc.err; cat $O/bench_tc.json; tail -3 $O/bench_tc.err
RNB_BACKWARD=mma timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_mma.json 2> $O/bench_mma.err; cat $O/bench_mma.json | cut -c1-300; python tools/bench_cmp.py $O
