#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
n=d.get("natural")
print(sys.argv[1], "ms/train", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "natural", (round(n["ms_per_step"],3), round(n["ms_resolve_per_train"],3), n["tie_sensitive_decisions"], n["repeated_with_compensated_sums"]) if n else None)
print("  centroids:", d["config"]["centroids"])
PY
}
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-cpp > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
echo "bench c2 exit $?"; summ gpurun_out/bench_c2_$TAG.json; tail -3 gpurun_out/bench_c2_$TAG.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2_r2j.json"))
print("sensitive per level:", d.get("sensitive_per_level"))
PY
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --exact --no-cpu --no-cpp > gpurun_out/bench_c2_exact_$TAG.json 2> gpurun_out/bench_c2_exact_$TAG.err
echo "bench c2 exact exit $?"; summ gpurun_out/bench_c2_exact_$TAG.json; tail -3 gpurun_out/bench_c2_exact_$TAG.err
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/bench_c3_exact_$TAG.json 2> gpurun_out/bench_c3_exact_$TAG.err
echo "bench c3 exact exit $?"; summ gpurun_out/bench_c3_exact_$TAG.json; tail -3 gpurun_out/bench_c3_exact_$TAG.err
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-cpp > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
echo "bench c3 auto exit $?"; summ gpurun_out/bench_c3_$TAG.json; tail -3 gpurun_out/bench_c3_$TAG.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --no-cpp > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
echo "bench c4 auto exit $?"; summ gpurun_out/bench_c4_$TAG.json; tail -3 gpurun_out/bench_c4_$TAG.err
QB200_TC_FUSE=1 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-cpp --no-natural --centroids integer > gpurun_out/bench_c2_fused_$TAG.json 2>/dev/null
python - <<PY
import json
for f in ("gpurun_out/bench_c2_fused_r2j.json",):
    d=json.load(open(f)); print(f, d["ms_per_step"], {k:v["ms_assign"] for k,v in d["per_level_ms"].items()})
PY
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-cpp --no-natural --centroids integer > gpurun_out/bench_c2_integer_$TAG.json 2>/dev/null
python - <<PY
import json
for f in ("gpurun_out/bench_c2_integer_r2j.json",):
    d=json.load(open(f)); print(f, d["ms_per_step"], {k:v["ms_assign"] for k,v in d["per_level_ms"].items()})
PY
timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-cpp --no-natural --centroids integer > gpurun_out/bench_c3_integer_$TAG.json 2>/dev/null
python - <<PY
import json
for f in ("gpurun_out/bench_c3_integer_r2j.json",):
    d=json.load(open(f)); print(f, d["ms_per_step"], {k:v["ms_assign"] for k,v in d["per_level_ms"].items()}, {k:v["ms_resolve"] for k,v in d["per_level_ms"].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_exact_$TAG.csv \
    python bench.py --workload c2 --steps 1 --warmup 3 --exact --no-cpu --no-cpp --no-natural > gpurun_out/ncu_launches_exact_$TAG.log 2>&1
echo "ncu exit $?"
