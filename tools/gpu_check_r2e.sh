#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fx_runs_kernel -s 40 -c 1 \
    -o gpurun_out/prof_fx_runs_$TAG -f python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/ncu_fx_runs_$TAG.log 2>&1
echo "ncu runs exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fx_chain_kernel -s 33 -c 1 \
    -o gpurun_out/prof_fx_chain_$TAG -f python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/ncu_fx_chain_$TAG.log 2>&1
echo "ncu chain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_exact_natural_$TAG.csv \
    python bench.py --workload c2 --data natural --steps 1 --warmup 3 --exact --no-cpu --no-cpp > gpurun_out/ncu_launches_exact_nat_$TAG.log 2>&1
echo "ncu natural exit $?"
