#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "default rc=$?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.time
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time; echo "reference rc=$?"; cat gpurun_out/bench_reference.json; cat gpurun_out/bench_reference.time
bash tools/gpu_ncu_gemm.sh "16384 2048 512 fwd 4" gemm2_conv1 | head -8
