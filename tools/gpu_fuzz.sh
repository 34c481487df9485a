#!/bin/bash
# fuzz only: default (auto) mode with several seeds, then the always-exact and integer modes
set -u
mkdir -p gpurun_out
TAG=${1:-r2fuzz}
for seed in 101 202 303; do
timeout 300 python tools/fuzz_parity.py 110 $seed auto 2>&1 | tail -4
done
timeout 200 python tools/fuzz_parity.py 60 404 exact 2>&1 | tail -3
timeout 200 python tools/fuzz_parity.py 50 505 2>&1 | tail -3
