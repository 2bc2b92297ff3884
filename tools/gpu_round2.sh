#!/bin/bash
# full GPU tests + bench + trace + ncu launch list of one eager step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/t_all.log)"
grep -E "^E  |FAILED" gpurun_out/t_all.log | head -20
timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>gpurun_out/bench_graph.err; cut -c1-200 gpurun_out/bench_graph.log
timeout -k 10 300 python tools/trace_step.py 2>&1 | tail -8
timeout -k 10 300 python tools/gemm_variants.py base 2>&1 | grep "^\["
