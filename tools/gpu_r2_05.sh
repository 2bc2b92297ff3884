#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python tools/debug_decoder.py > gpurun_out/r2_05_debug_decoder.log 2>&1; tail -12 gpurun_out/r2_05_debug_decoder.log | cut -c1-200
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests/test_gpu_parity.py -k "trainer_bound" -s 2>&1 | grep -E "^E  |passed|failed" | cut -c1-300 | head
for SMS in "0,0" "48,100" "40,108" "56,92"; do
  SAVQA_BRANCH_SMS=$SMS timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_05_bench_$SMS.json 2> gpurun_out/r2_05_bench.err; echo "bench branch_sms=$SMS rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2_05_bench_$SMS.json')); print({k: d[k] for k in ('value','ms_per_step','loss')}, d['e2e']['value'])" 2>&1 | tail -1
done
