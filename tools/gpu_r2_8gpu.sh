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
d=[json.loads(l) for l in open('gpurun_out/r2_bench_8gpu_$name.json') if l.startswith('{')][-1]; print('$name', {k: round(d[k],3) for k in ('value','ms_per_step','loss')}, 'e2e', round(d['e2e']['value']), d.get('ranks',{}).get('param_checksums_equal'))" 2>&1 | tail -1
}
VARIANTS=${VARIANTS:-"default nvls ll128 buckets3"}
for v in $VARIANTS; do
  case $v in
    default) run default SAVQA_X=1 ;;
    info) run info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING; grep -E "NVLS|Connected|Channel|algo|proto" gpurun_out/r2_bench_8gpu_info.err | cut -c1-200 | sort | uniq -c | sort -rn | head -30 ;;
    conn8) run conn8 CUDA_DEVICE_MAX_CONNECTIONS=8 ;;
    adam) run adam SAVQA_ADAM_PER_BUCKET=1 ;;
    adamb3) run adamb3 SAVQA_ADAM_PER_BUCKET=1 SAVQA_BUCKET_BLOCKS=3 ;;
    adamb2) run adamb2 SAVQA_ADAM_PER_BUCKET=1 SAVQA_BUCKET_BLOCKS=2 ;;
    nvls) run nvls NCCL_ALGO=NVLS ;;
    tree) run tree NCCL_ALGO=Tree ;;
    ll128) run ll128 NCCL_PROTO=LL128 ;;
    simple) run simple NCCL_PROTO=Simple ;;
    buckets3) run buckets3 SAVQA_BUCKET_BLOCKS=3 ;;
    buckets2) run buckets2 SAVQA_BUCKET_BLOCKS=2 ;;
    lazytables) run lazytables SAVQA_TABLES_EAGER=0 ;;
    onecomm) run onecomm SAVQA_TABLE_COMM=0 ;;
    reserve12) run reserve12 SAVQA_BRANCH_SMS=52,84 NCCL_MAX_CTAS=12 NCCL_MAX_NCHANNELS=12 ;;
  esac
done
env ${TRACE_ENV:-SAVQA_X=1} timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/trace_step.py > gpurun_out/r2_trace_8gpu.log 2>&1
tail -8 gpurun_out/r2_trace_8gpu.log; cp gpurun_out/trace_step.json.gz gpurun_out/trace_step_8gpu.json.gz
python tools/summarize_trace.py gpurun_out/trace_step_8gpu.json.gz > gpurun_out/r2_trace_8gpu_summary.txt 2>&1; head -50 gpurun_out/r2_trace_8gpu_summary.txt
