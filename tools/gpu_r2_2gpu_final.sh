#!/bin/bash
# 2 GPUs, final code: data-parallel equality check (two of the four cases) + the 2-GPU bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for CASE in graph:full eager:encoder; do
  MODE=${CASE%%:*}; STEP=${CASE##*:}
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_check.py --mode $MODE --step $STEP --steps 6 > gpurun_out/r2_dp_check_${MODE}_$STEP.log 2>&1
  echo "dp_check $MODE $STEP rc=$?"; grep DP_CHECK gpurun_out/r2_dp_check_${MODE}_$STEP.log | cut -c1-600
done
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench 2gpu rc=$?"; cut -c1-260 gpurun_out/r2_bench_2gpu.json; grep -o '"ranks": {.*"param_checksums_equal": [a-z]*' gpurun_out/r2_bench_2gpu.json | cut -c1-80; grep -o '"e2e": {[^}]*}' gpurun_out/r2_bench_2gpu.json
