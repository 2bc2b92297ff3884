#!/bin/bash
# 8 GPUs: scaling bench (default, merged buckets, NCCL protocol), kernel timeline of rank 0
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu_$name.json 2> gpurun_out/r2_bench_8gpu_$name.err
  echo "bench 8gpu $name rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench_8gpu_$name.json')); print('$name', {k: round(d[k],3) for k in ('value','ms_per_step','loss')}, 'e2e', round(d['e2e']['value']), d.get('ranks',{}).get('param_checksums_equal'))" 2>&1 | tail -1
}
run default SAVQA_X=1
run reserve12 SAVQA_BRANCH_SMS=52,84 NCCL_MAX_CTAS=12 NCCL_MAX_NCHANNELS=12
run buckets3 SAVQA_BUCKET_BLOCKS=3
run simple NCCL_PROTO=Simple
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/trace_step.py > gpurun_out/r2_trace_8gpu.log 2>&1
tail -8 gpurun_out/r2_trace_8gpu.log; cp gpurun_out/trace_step.json.gz gpurun_out/trace_step_8gpu.json.gz
python tools/summarize_trace.py gpurun_out/trace_step_8gpu.json.gz > gpurun_out/r2_trace_8gpu_summary.txt 2>&1; head -50 gpurun_out/r2_trace_8gpu_summary.txt
