#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests/test_gpu_parity.py > gpurun_out/t_par.log 2>&1; echo "parity rc=$? $(tail -n 1 gpurun_out/t_par.log)"
grep -E "^E  |FAILED" gpurun_out/t_par.log | head
for i in 1 2; do timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*'; done
