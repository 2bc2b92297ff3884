#!/bin/bash
# quick confirmation after a change: the whole GPU suite, LayerNorm microbench, graph bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/t_all.log)"
grep -E "^E  |FAILED" gpurun_out/t_all.log | head
timeout -k 10 120 python tools/bench_kernels.py ln 2>&1 | cut -c1-140
for i in 1 2; do timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*'; done
