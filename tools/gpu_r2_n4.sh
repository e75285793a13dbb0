#!/bin/bash
O=gpurun_out/r2dpfinal_n4; mkdir -p $O
bash tools/gpu_canary.sh 150 || exit 1
export RNB_BENCH_CACHE=/dev/shm/rnb_bench_cache
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 --steps 200 --warmup 20 --no-records > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench.json")); print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["roofline"]["stages"].items()})
PY
