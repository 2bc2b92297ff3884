#!/bin/bash
# single GPU: early merged Adam on/off, full GPU suite, kernel timeline
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests/test_gpu_parity.py > gpurun_out/r2_21_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_21_pytest.log)"
grep -E "^E  |FAILED" gpurun_out/r2_21_pytest.log | cut -c1-400 | head -20
for v in 1 0 1 0; do
  SAVQA_EARLY_ADAM=$v timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_21_bench_early$v.json 2> gpurun_out/r2_21_bench_early$v.err; echo "early_adam=$v rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2_21_bench_early$v.json')); print({k: round(d[k],3) for k in ('value','ms_per_step','loss')}, round(d['e2e']['value']))"
done
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_21_trace.log 2>&1; tail -3 gpurun_out/r2_21_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_21_trace_summary.txt 2>&1; sed -n 1,16p gpurun_out/r2_21_trace_summary.txt
