#!/bin/bash
# full GPU suite + mesh timing + snapshot pin in one call
O=gpurun_out/${1:-r2}; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log | cut -c1-300
timeout 400 python tests/ref_pin_snapshot.py --out $O > $O/pin_snapshot.log 2>&1; echo "rc=$?" >> $O/pin_snapshot.log; python - <<PY
import json
d=json.load(open("$O/summary_snapshot.json")); print(d["reencode_byte_identical"], d["A"]["movement_defaults_equal"], d["B"]["probe_cuda_vs_ref"], d["B"].get("probe_cuda_all_levels_vs_ref"))
PY
timeout 400 python tools/mesh_time.py $O/mesh_time.json > $O/mesh_time.log 2>&1; echo "rc=$?" >> $O/mesh_time.log; tail -4 $O/mesh_time.log | cut -c1-700
