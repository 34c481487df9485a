#!/bin/bash
# multi-GPU checks: usage tools/gpu_check_multi.sh TAG NGPU
set -u
mkdir -p gpurun_out
TAG=${1:-r2m}; N=${2:-2}
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
timeout 300 quant_b200/host/host_test multi 4096 4096 2 2 10 $N 0 2>&1 | tail -3
timeout 300 quant_b200/host/host_test multi 4096 4096 2 2 10 $N 1 2>&1 | tail -3
timeout 600 quant_b200/host/host_test multi 16384 16384 2 2 12 $N 0 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_device or sharded or two_rank" > gpurun_out/pytest_multi_$TAG.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_multi_$TAG.log
for mode in "" "--nccl"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-natural $mode > gpurun_out/bench_c3_n${N}${mode}_$TAG.json 2> gpurun_out/bench_c3_n${N}${mode}_$TAG.err
echo "bench c3 N=$N $mode exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c3_n${N}${mode}_$TAG.json"))
    print("N=$N $mode ms/train", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["config"].get("allreduce"))
except Exception as e: print("no json", e)
PY
tail -4 gpurun_out/bench_c3_n${N}${mode}_$TAG.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c2 --weak --steps 10 --warmup 3 --no-natural > gpurun_out/bench_c2weak_n${N}_$TAG.json 2> gpurun_out/bench_c2weak_n${N}_$TAG.err
echo "bench c2 weak exit $?"; cat gpurun_out/bench_c2weak_n${N}_$TAG.json | cut -c1-400
