#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x"
timeout -k 10 400 $PT tests/test_gpu_kernels.py -k "attention_forward" > gpurun_out/a_fwd.log 2>&1; echo "fwd rc=$?"; tail -n 12 gpurun_out/a_fwd.log
timeout -k 10 400 $PT tests/test_gpu_kernels.py -k "attention_backward" > gpurun_out/a_bwd.log 2>&1; echo "bwd rc=$?"; tail -n 12 gpurun_out/a_bwd.log
timeout -k 10 400 $PT tests/test_gpu_parity.py > gpurun_out/a_par.log 2>&1; echo "parity rc=$?"; tail -n 12 gpurun_out/a_par.log
timeout -k 10 300 python tools/bench_kernels.py attn > gpurun_out/kbench_attn.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kbench_attn.log
timeout -k 10 300 python bench.py --mode graph --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_graph.log
