#!/bin/bash
# small-cell sums (auto mode): tests, fuzz in the default mode, bench lines of every workload
set -u
mkdir -p gpurun_out
TAG=${1:-r2small}
timeout 900 python -m pytest tests -m gpu -x -q -k "small_cells or multi_device or auto_mode or full_train or duplicate_heavy or two_rank or sharded" > gpurun_out/pytest_small_$TAG.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_small_$TAG.log
timeout 200 python tools/fuzz_parity.py 100 77 auto 2>&1 | tail -4
timeout 900 python bench.py --steps 5 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
echo "bench c3 exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c3_$TAG.json"))
print("c3 ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "centroids", d["config"].get("centroids"), "sens", d.get("sensitive_per_level"), "nat", d.get("natural"))
PY
for wl in c2 c4 c1; do
timeout 600 python bench.py --workload $wl --no-cpu > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
echo "bench $wl exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_$TAG.json"))
print("$wl ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "centroids", d["config"].get("centroids"), "sens", d.get("sensitive_per_level"), "nat", d.get("natural"))
PY
done
