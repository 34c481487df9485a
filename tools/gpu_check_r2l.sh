#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2l}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
n=d.get("natural")
print(sys.argv[1], "ms/train", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "natural", (round(n["ms_per_step"],3), round(n["ms_resolve_per_train"],3), n["tie_sensitive_decisions"], n["repeated_with_compensated_sums"]) if n else None)
print("  centroids:", d["config"]["centroids"], d.get("sensitive_per_level"), d.get("integer_sum_mode"))
PY
}
for wl in c3 c1 c4; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-cpp > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
echo "bench $wl exit $?"; summ gpurun_out/bench_${wl}_$TAG.json; tail -3 gpurun_out/bench_${wl}_$TAG.err
done
timeout 300 python tools/fuzz_parity.py 150 94 auto 2>&1 | tail -3
