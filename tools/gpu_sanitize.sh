#!/bin/bash
# compute-sanitizer passes (the reference has none) on small workloads; logs -> gpurun_out/<tag>/sanitizer_*.log
O=gpurun_out/${1:-san}; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py ${2:-all} > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> $O/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case done|rc=" $O/sanitizer_$tool.log | tail -4
done
