#!/bin/bash
# 2 GPUs: scheduling variants of the table exchange / buckets + kernel timeline of rank 0
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu_$name.json 2> gpurun_out/r2_bench_2gpu_$name.err
  echo "bench 2gpu $name rc=$?"
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_2gpu_$name.json') if l.startswith('{')][-1]; print('$name', {k: round(d[k],3) for k in ('value','ms_per_step','loss')}, 'e2e', round(d['e2e']['value']), d.get('ranks',{}).get('param_checksums_equal'))" 2>&1 | tail -1
}
for v in ${VARIANTS:-default lazytables onecomm buckets3}; do
  case $v in
    default) run default SAVQA_X=1 ;;
    lazytables) run lazytables SAVQA_TABLES_EAGER=0 ;;
    onecomm) run onecomm SAVQA_TABLE_COMM=0 ;;
    lazyone) run lazyone SAVQA_TABLE_COMM=0 SAVQA_TABLES_EAGER=0 ;;
    buckets3) run buckets3 SAVQA_BUCKET_BLOCKS=3 ;;
    buckets2) run buckets2 SAVQA_BUCKET_BLOCKS=2 ;;
    conn32) run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 ;;
    conn32lazy) run conn32lazy CUDA_DEVICE_MAX_CONNECTIONS=32 SAVQA_TABLES_EAGER=0 ;;
    conn32b3) run conn32b3 CUDA_DEVICE_MAX_CONNECTIONS=32 SAVQA_BUCKET_BLOCKS=3 ;;
    c32adam) run c32adam CUDA_DEVICE_MAX_CONNECTIONS=32 SAVQA_ADAM_PER_BUCKET=1 ;;
    c32adamb3) run c32adamb3 CUDA_DEVICE_MAX_CONNECTIONS=32 SAVQA_ADAM_PER_BUCKET=1 SAVQA_BUCKET_BLOCKS=3 ;;
    c32b2) run c32b2 CUDA_DEVICE_MAX_CONNECTIONS=32 SAVQA_BUCKET_BLOCKS=2 ;;
    lazyb3) run lazyb3 SAVQA_TABLES_EAGER=0 SAVQA_BUCKET_BLOCKS=3 ;;
  esac
done
for t in ${TRACES:-default}; do
  case $t in
    default) E="SAVQA_X=1" ;;
    lazytables) E="SAVQA_TABLES_EAGER=0" ;;
    onecomm) E="SAVQA_TABLE_COMM=0" ;;
    conn32) E="CUDA_DEVICE_MAX_CONNECTIONS=32" ;;
    c32adam) E="CUDA_DEVICE_MAX_CONNECTIONS=32 SAVQA_ADAM_PER_BUCKET=1" ;;
  esac
  env $E timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 tools/trace_step.py > gpurun_out/r2_trace_2gpu_$t.log 2>&1
  tail -3 gpurun_out/r2_trace_2gpu_$t.log; cp gpurun_out/trace_step.json.gz gpurun_out/trace_step_2gpu_$t.json.gz
  python - <<PY
import gzip, json
rows=json.load(gzip.open('gpurun_out/trace_step_2gpu_$t.json.gz','rt'))
print('trace $t span', max(r['t']+r['d'] for r in rows))
for r in sorted(rows,key=lambda r:r['t']):
    if 'nccl' in r['n']: print(f"{r['s']:4d} {r['t']:8.1f} {r['d']:6.1f} {r['t']+r['d']:8.1f} {r['n'][:40]}")
PY
done
