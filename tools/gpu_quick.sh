#!/bin/bash
# quick confirmation after a host-side change: parity tests + graph bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x"
timeout -k 10 400 $PT tests/test_gpu_parity.py > gpurun_out/q_par.log 2>&1; echo "parity rc=$?"; tail -n 4 gpurun_out/q_par.log
timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>gpurun_out/bench_graph.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_graph.log; tail -3 gpurun_out/bench_graph.err
