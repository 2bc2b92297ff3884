#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x"
timeout -k 10 600 $PT tests/test_gpu_kernels.py > gpurun_out/p_k.log 2>&1; echo "kernels rc=$?"; tail -n 3 gpurun_out/p_k.log
timeout -k 10 400 $PT tests/test_gpu_parity.py > gpurun_out/p_par.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/p_par.log
timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl1.log 2>gpurun_out/bench_pdl1.err; echo "bench pdl rc=$?"; cut -c1-230 gpurun_out/bench_pdl1.log; tail -2 gpurun_out/bench_pdl1.err
SAVQA_PDL=0 timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl0.log 2>gpurun_out/bench_pdl0.err; echo "bench nopdl rc=$?"; cut -c1-230 gpurun_out/bench_pdl0.log
timeout -k 10 200 python tools/chain_bench.py 2>&1 | head -8
SAVQA_PDL=0 timeout -k 10 200 python tools/chain_bench.py 2>&1 | head -8
