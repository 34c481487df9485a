#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "regressions or auto_mode or small_cells or full_train or duplicate_heavy or end_to_end" > gpurun_out/pytest_last.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_last.log
timeout 120 python tools/fuzz_parity.py 70 909 auto 2>&1 | tail -3
