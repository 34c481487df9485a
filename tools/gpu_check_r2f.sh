#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
n=d.get("natural")
print(sys.argv[1], "ms/train", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "natural", (round(n["ms_per_step"],3), round(n["ms_resolve_per_train"],3), n["tie_sensitive_decisions"], n["repeated_with_compensated_sums"]) if n else None)
print("  assign", {k:v["ms_assign"] for k,v in d["per_level_ms"].items()})
print("  resolve", {k:v["ms_resolve"] for k,v in d["per_level_ms"].items()}, "flagged last", d["flagged_last_level"])
print("  centroids:", d["config"]["centroids"])
PY
}
for wl in c2 c3; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-cpp > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
echo "bench $wl exit $?"; summ gpurun_out/bench_${wl}_$TAG.json; tail -3 gpurun_out/bench_${wl}_$TAG.err
timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --exact --no-cpu --no-cpp > gpurun_out/bench_${wl}_exact_$TAG.json 2> gpurun_out/bench_${wl}_exact_$TAG.err
echo "bench $wl exact exit $?"; summ gpurun_out/bench_${wl}_exact_$TAG.json; tail -3 gpurun_out/bench_${wl}_exact_$TAG.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_exact_$TAG.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp > gpurun_out/ncu_launches_exact_$TAG.log 2>&1
echo "ncu exit $?"
