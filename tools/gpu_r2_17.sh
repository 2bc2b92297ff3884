#!/bin/bash
set -u
cd "$(dirname "$0")/.."
VARIANTS="conn32 conn32lazy conn32b3 lazyb3" TRACES="conn32" bash tools/gpu_r2_2gpu_b.sh
for c in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_17_bench_conn$c.json 2> gpurun_out/r2_17_bench_conn$c.err; echo "1gpu conn$c rc=$?"; cut -c1-230 gpurun_out/r2_17_bench_conn$c.json
done
