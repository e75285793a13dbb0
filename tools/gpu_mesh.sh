#!/bin/bash
# On the B200 box: mesh path (N2) GPU tests, reference pin of the mesh path, timing
O=gpurun_out/${1:-mesh}; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_mesh.py -x -q > $O/pytest_mesh.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mesh.log; tail -30 $O/pytest_mesh.log | cut -c1-400
timeout 400 python tests/ref_pin_mesh.py --config full --steps 120 --rays 4096 --mesh 30 --golden lattice32 --out $O > $O/pin_mesh_32.log 2>&1; echo "rc=$?" >> $O/pin_mesh_32.log; tail -3 $O/pin_mesh_32.log | cut -c1-1500
timeout 500 python tests/ref_pin_mesh.py --config full --steps 120 --rays 4096 --mesh 256 --out $O > $O/pin_mesh_full.log 2>&1; echo "rc=$?" >> $O/pin_mesh_full.log; tail -3 $O/pin_mesh_full.log | cut -c1-1500
if [ "$2" == "time" ]; then timeout 400 python tools/mesh_time.py $O/mesh_time.json > $O/mesh_time.log 2>&1; echo "rc=$?" >> $O/mesh_time.log; tail -5 $O/mesh_time.log | cut -c1-400; fi
if [ "$3" == "ncu" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:'k_mc_|k_scan_counts|k_mesh_' --csv --log-file $O/mesh_ncu_512.csv python tools/mesh_ncu.py 512 > $O/mesh_ncu.log 2>&1; echo "ncu rc=$?" >> $O/mesh_ncu.log; tail -2 $O/mesh_ncu.log | cut -c1-300
fi
if [ "$4" == "snap" ]; then
  timeout 400 python tests/ref_pin_snapshot.py --out $O > $O/pin_snapshot.log 2>&1; echo "rc=$?" >> $O/pin_snapshot.log; tail -4 $O/pin_snapshot.log | cut -c1-3000
fi
