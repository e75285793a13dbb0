#!/bin/bash
# first thing in every GPU session: the network kernels on a tiny batch under a short timeout.  A deadlocked kernel must cost seconds, not the session.
timeout ${1:-150} python tools/sanitize_case.py network > gpurun_out/canary.log 2>&1; rc=$?
tail -3 gpurun_out/canary.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "CANARY FAILED rc=$rc: not running the rest of the session"; exit 1; fi
