#!/bin/bash
# final single-GPU measurement set of round 2: tests, bench lines for every workload, reference arm, launch lists, ncu captures
set -u
mkdir -p gpurun_out
TAG=${1:-r2final}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
echo "bench default (c3) exit $?"; cut -c1-600 gpurun_out/bench_c3_$TAG.json; tail -3 gpurun_out/bench_c3_$TAG.err
for wl in c2 c1 c4; do
timeout 600 python bench.py --workload $wl > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
echo "bench $wl exit $?"; cut -c1-300 gpurun_out/bench_${wl}_$TAG.json
done
timeout 600 python bench.py --workload c5 --steps 5 > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err
echo "bench c5 exit $?"; cut -c1-300 gpurun_out/bench_c5_$TAG.json
for wl in c2 c3; do
timeout 600 python bench.py --workload $wl --steps 3 --exact --no-cpu --no-cpp > gpurun_out/bench_${wl}_exact_$TAG.json 2> gpurun_out/bench_${wl}_exact_$TAG.err
echo "bench $wl exact exit $?"; cut -c1-300 gpurun_out/bench_${wl}_exact_$TAG.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference_$TAG.json 2> gpurun_out/bench_c3_reference_$TAG.err
echo "reference arm exit $?"; cut -c1-600 gpurun_out/bench_c3_reference_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural --centroids integer > gpurun_out/ncu_l3_$TAG.log 2>&1
echo "ncu launches c3 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_$TAG.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_l2_$TAG.log 2>&1
echo "ncu launches c2 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_exact_$TAG.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/ncu_l2e_$TAG.log 2>&1
echo "ncu launches c2 exact exit $?"
# full captures: the K = 4096 tensor-core launch of config 3 (second pass of the last level), K = 1024 of config 2, re-rank, exact-sum kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assign_tc_kernel -s 27 -c 1 -o gpurun_out/prof_tc_c3_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural --centroids integer > gpurun_out/ncu_p1_$TAG.log 2>&1
echo "ncu tc c3 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assign_tc_kernel -s 15 -c 1 -o gpurun_out/prof_tc_c2_$TAG -f \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_p2_$TAG.log 2>&1
echo "ncu tc c2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:refilter_kernel -s 15 -c 1 -o gpurun_out/prof_refilter_c2_$TAG -f \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_p3_$TAG.log 2>&1
echo "ncu refilter exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fx_runs_kernel -s 40 -c 1 -o gpurun_out/prof_fx_runs_$TAG -f \
    python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/ncu_p4_$TAG.log 2>&1
echo "ncu fx_runs exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_smem -s 15 -c 1 -o gpurun_out/prof_acc_c2_$TAG -f \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_p5_$TAG.log 2>&1
echo "ncu accumulate exit $?"
ls -la gpurun_out | grep $TAG | head -40
