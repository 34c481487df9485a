#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, a short bench, the ncu launch list and one full capture
# of the assignment and accumulate kernels.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ "${NCU:-1}" = "1" ]; then
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:assign_tc_kernel -s 3 -c 1 \
    -o gpurun_out/prof_assign_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_assign_$TAG.log 2>&1
echo "ncu assign exit $?"
ncu --set full --clock-control none --import-source on -k regex:accumulate_smem -s 3 -c 1 \
    -o gpurun_out/prof_acc_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_acc_$TAG.log 2>&1
echo "ncu acc exit $?"
fi
ls -la gpurun_out
