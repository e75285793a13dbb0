#!/bin/bash
# round 2, second GPU session: full GPU suite, then the stage-2 resume probe (reference vs library from the same snapshot)
O=gpurun_out/${1:-r2b}; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -30 $O/pytest_gpu.log | cut -c1-400
timeout 900 python tools/resume_probe.py $O/resume > $O/resume.log 2>&1; echo "resume rc=$?" >> $O/resume.log; tail -12 $O/resume.log | cut -c1-900
