#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests -s > gpurun_out/r2_04_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_04_pytest.log)"
grep -E "^E  |FAILED|fused decoder vs chain" gpurun_out/r2_04_pytest.log | cut -c1-1500 | head -30
timeout -k 10 600 python tools/debug_decoder.py > gpurun_out/r2_04_debug_decoder.log 2>&1; tail -52 gpurun_out/r2_04_debug_decoder.log | cut -c1-200
timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_04_bench.json 2> gpurun_out/r2_04_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_04_bench.json')); print({k: d[k] for k in ('value','ms_per_step','loss','gpu_launches')}, d['e2e']['value'])" 2>&1 | tail -1
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_04_trace.log 2>&1; tail -3 gpurun_out/r2_04_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_04_trace_summary.txt 2>&1; head -44 gpurun_out/r2_04_trace_summary.txt
