#!/bin/bash
set -u
mkdir -p gpurun_out
QB200_DEBUG_TREE=1 timeout 600 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural --centroids integer > gpurun_out/bench_c3_diag.json 2> gpurun_out/bench_c3_diag.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c3_diag.json"))
print(d["ms_per_step"], d["sensitive_per_level"], d.get("sensitive_diag"))
PY
grep "robustness" gpurun_out/bench_c3_diag.err | tail -14
