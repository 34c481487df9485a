#!/bin/bash
# 2-GPU box: regression tests of the census fixes + the N=2 strong-scaling line of config 3
set -u
mkdir -p gpurun_out
TAG=${1:-r2n2}
timeout 400 python -m pytest tests -m gpu -q -k "regressions or auto_mode or small_cells or multi_device" > gpurun_out/pytest_n2_$TAG.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_n2_$TAG.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-natural > gpurun_out/bench_c3_n2_$TAG.json 2> gpurun_out/bench_c3_n2_$TAG.err
echo "bench c3 N=2 exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c3_n2_$TAG.json"))
    print("N=2 ms/train", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["config"].get("centroids"))
except Exception as e: print("no json", e)
PY
tail -3 gpurun_out/bench_c3_n2_$TAG.err
