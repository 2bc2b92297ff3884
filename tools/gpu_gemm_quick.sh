#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x"
timeout -k 10 400 $PT tests/test_gpu_kernels.py -k "gemm or layernorm" > gpurun_out/g_all.log 2>&1; echo "gemm+ln tests rc=$?"; tail -n 3 gpurun_out/g_all.log
timeout -k 10 300 python tools/bench_kernels.py gemm ln 2>&1 | cut -c1-120 | grep -v "dec \|head"
timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-230
