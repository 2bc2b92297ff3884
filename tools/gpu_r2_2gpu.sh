#!/bin/bash
# 2 GPUs: on-hardware data-parallel equality check (tests/dp_check.py, graph + eager), 2-GPU bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for CASE in graph:encoder eager:encoder graph:full eager:full; do
  MODE=${CASE%%:*}; STEP=${CASE##*:}
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_check.py --mode $MODE --step $STEP --steps 6 > gpurun_out/r2_dp_check_${MODE}_$STEP.log 2>&1
  echo "dp_check $MODE $STEP rc=$?"; grep DP_CHECK gpurun_out/r2_dp_check_${MODE}_$STEP.log | cut -c1-900; grep -E "Error|error" gpurun_out/r2_dp_check_${MODE}_$STEP.log | head -5
done
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench 2gpu rc=$?"; cut -c1-300 gpurun_out/r2_bench_2gpu.json; grep -o '"ranks": {[^}]*}' gpurun_out/r2_bench_2gpu.json | cut -c1-300; grep -o '"e2e": {[^}]*}' gpurun_out/r2_bench_2gpu.json
