#!/bin/bash
# last single-GPU run of round 2: full tests, fuzz in the default mode, the bench lines, one launch list
set -u
mkdir -p gpurun_out
TAG=${1:-r2final2}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/fuzz_parity.py 170 78 auto 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
echo "bench c3 exit $?"; tail -3 gpurun_out/bench_c3_$TAG.err
for wl in c2 c4 c1; do
timeout 600 python bench.py --workload $wl > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
echo "bench $wl exit $?"
done
python - <<PY
import json
for wl in ("c3","c2","c4","c1"):
    try:
        d=json.load(open(f"gpurun_out/bench_{wl}_$TAG.json"))
        print(wl, "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "cpp", (d.get("e2e_cpp") or {}).get("seconds_per_train"),
              "sens", sum((d.get("sensitive_per_level") or {}).values()), "nat", (d.get("natural") or {}).get("ms_per_step"), (d.get("natural") or {}).get("tie_sensitive_decisions"))
    except Exception as e: print(wl, "no json", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-cpp --no-natural > gpurun_out/ncu_l3_$TAG.log 2>&1
echo "ncu launches c3 exit $?"
