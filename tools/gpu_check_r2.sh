#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, bench lines, the ncu launch list and full captures.
# usage: tools/gpu_check_r2.sh TAG [tests|notests] [ncu|noncu]
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
if [ "${2:-tests}" = "tests" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
echo "bench c3 exit $?"; cat gpurun_out/bench_c3_$TAG.json; tail -5 gpurun_out/bench_c3_$TAG.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
echo "bench c2 exit $?"; cat gpurun_out/bench_c2_$TAG.json; tail -5 gpurun_out/bench_c2_$TAG.err
if [ "${3:-ncu}" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_$TAG.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assign_tc_kernel -s 15 -c 1 \
    -o gpurun_out/prof_assign_c2_$TAG -f python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_assign_$TAG.log 2>&1
echo "ncu assign exit $?"
fi
ls -la gpurun_out | tail -20
