#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests -s > gpurun_out/r2_06_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_06_pytest.log)"
grep -E "^E  |FAILED|bound vs autograd" gpurun_out/r2_06_pytest.log | cut -c1-600 | head -20
run() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_06_bench_$name.json 2> gpurun_out/r2_06_bench.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_06_bench_$name.json')); print('$name', {k: round(d[k],3) for k in ('value','ms_per_step','loss')}, round(d['e2e']['value']))" 2>&1 | tail -1
}
run default SAVQA_X=1
run b52_96 SAVQA_BRANCH_SMS=52,96
run b60_88 SAVQA_BRANCH_SMS=60,88
run b64_84 SAVQA_BRANCH_SMS=64,84
run b48_92 SAVQA_BRANCH_SMS=48,92
run side148 SAVQA_SIDE_SMS=148
run side100 SAVQA_SIDE_SMS=100
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_06_trace.log 2>&1; tail -3 gpurun_out/r2_06_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_06_trace_summary.txt 2>&1; head -44 gpurun_out/r2_06_trace_summary.txt
