#!/bin/bash
# On the B200 box: full GPU suite, then optional extras: snap (snapshot pin), mesh (mesh timing), dataset (ingest timing)
O=gpurun_out/${1:-r2}; mkdir -p $O; shift
timeout 700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log | cut -c1-300
for what in "$@"; do
  case $what in
    snap) timeout 400 python tests/ref_pin_snapshot.py --out $O > $O/pin_snapshot.log 2>&1; echo "rc=$?" >> $O/pin_snapshot.log; tail -2 $O/pin_snapshot.log | cut -c1-300;;
    mesh) timeout 400 python tools/mesh_time.py $O/mesh_time.json > $O/mesh_time.log 2>&1; echo "rc=$?" >> $O/mesh_time.log; tail -4 $O/mesh_time.log | cut -c1-700;;
    dataset) timeout 500 python tools/dataset_time.py $O/dataset_time.json > $O/dataset_time.log 2>&1; echo "rc=$?" >> $O/dataset_time.log; tail -2 $O/dataset_time.log | cut -c1-900;;
  esac
done
