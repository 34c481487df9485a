#!/bin/bash
# second-round helper: targeted tests + exact-mode bench lines
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q -k "multi or exact or parallel or two_rank or sharded" > gpurun_out/pytest_sel_$TAG.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_sel_$TAG.log
timeout 300 quant_b200/host/host_test multi 4096 4096 2 2 10 1 0 2>&1 | tail -3
timeout 300 quant_b200/host/host_test multi 4096 4096 2 2 10 1 1 2>&1 | tail -3
for wl in c2 c3; do
timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --exact --no-cpu --no-cpp > gpurun_out/bench_${wl}_exact_$TAG.json 2> gpurun_out/bench_${wl}_exact_$TAG.err
echo "bench $wl exact exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_exact_$TAG.json"))
print("$wl exact ms/train", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "natural", d["natural"]["ms_per_step"] if d.get("natural") else None)
PY
tail -3 gpurun_out/bench_${wl}_exact_$TAG.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_exact_$TAG.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/ncu_launches_exact_$TAG.log 2>&1
echo "ncu exit $?"
