#!/bin/bash
# round 2, first GPU session: full GPU suite (FULL-size parity, controller, drop-in test), then the drop-in record of both binaries
O=gpurun_out/${1:-r2a}; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dropin.py > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log | cut -c1-400
timeout 700 python -m pytest tests/test_gpu_dropin.py -m gpu -q > $O/pytest_dropin.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dropin.log; tail -25 $O/pytest_dropin.log | cut -c1-400
timeout 1500 python tools/dropin_run.py $O/dropin --iters ${2:-3000} --res ${3:-256} --pipeline > $O/dropin.log 2>&1; echo "dropin rc=$?" >> $O/dropin.log; tail -12 $O/dropin.log | cut -c1-600
# keep the record and the logs, not the scenes / meshes
rm -rf $O/dropin/scene_* $O/dropin/rnb_input $O/dropin/pipeline_*/prepared_data
ls -la $O/dropin | head -30
