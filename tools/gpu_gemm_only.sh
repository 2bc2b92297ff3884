#!/bin/bash
# GEMM bring-up: pair-kernel tests in their own processes, then the GEMM microbenchmark with both kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x"
timeout -k 10 300 $PT tests/test_gpu_kernels.py -k "pair_kernel" > gpurun_out/g_pair.log 2>&1; echo "pair rc=$?"; tail -n 15 gpurun_out/g_pair.log
timeout -k 10 300 $PT tests/test_gpu_kernels.py -k "dgrad" > gpurun_out/g_dgrad.log 2>&1; echo "dgrad rc=$?"; tail -n 8 gpurun_out/g_dgrad.log
timeout -k 10 300 $PT tests/test_gpu_kernels.py -k "wgrad" > gpurun_out/g_wgrad.log 2>&1; echo "wgrad rc=$?"; tail -n 8 gpurun_out/g_wgrad.log
timeout -k 10 300 python tools/bench_kernels.py gemm > gpurun_out/kbench_gemm2.log 2>&1; echo "kbench2 rc=$?"
SAVQA_GEMM_ENGINE=1 timeout -k 10 300 python tools/bench_kernels.py gemm > gpurun_out/kbench_gemm1.log 2>&1; echo "kbench1 rc=$?"
paste -d'\n' gpurun_out/kbench_gemm2.log gpurun_out/kbench_gemm1.log | cut -c1-150
