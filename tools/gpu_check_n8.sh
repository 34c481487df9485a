#!/bin/bash
# 8-GPU box, kept short (charged 8x): the 8-rank peer all-reduce / small-cell exchange tests and one strong-scaling bench line
set -u
mkdir -p gpurun_out
TAG=${1:-r2n8}; N=${2:-8}
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests -m gpu -q -k "multi_device and (8- or 3-auto or 4-1 or uneven)" > gpurun_out/pytest_multi_$TAG.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_multi_$TAG.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-natural > gpurun_out/bench_c3_n${N}_$TAG.json 2> gpurun_out/bench_c3_n${N}_$TAG.err
echo "bench c3 N=$N exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c3_n${N}_$TAG.json"))
    print("N=$N ms/train", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["config"].get("allreduce"), d["config"].get("centroids"))
    print("per level", d.get("per_level_ms"))
except Exception as e: print("no json", e)
PY
tail -4 gpurun_out/bench_c3_n${N}_$TAG.err
timeout 200 quant_b200/host/host_test multi 8192 8192 2 2 11 $N 0 2>&1 | tail -3
