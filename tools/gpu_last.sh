#!/bin/bash
# last short check of the round (budget: under 3 minutes)
set -u
mkdir -p gpurun_out
timeout 75 python -m pytest tests -m gpu -q -x -k "regressions or auto_mode or small_cells or duplicate_heavy" > gpurun_out/pytest_last.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_last.log
timeout 70 python bench.py --steps 3 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/bench_c3_last.json 2> gpurun_out/bench_c3_last.err
echo "bench exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c3_last.json")); print("c3 ms", d["ms_per_step"], "sens", d.get("sensitive_per_level"), d["config"]["centroids"][:90])
except Exception as e: print("no json", e)
PY
timeout 40 python tools/fuzz_parity.py 25 1001 auto 2>&1 | tail -2
