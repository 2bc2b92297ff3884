#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests > gpurun_out/r2_08_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_08_pytest.log)"
grep -E "^E  |FAILED" gpurun_out/r2_08_pytest.log | cut -c1-400 | head -20
timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_08_bench.json 2> gpurun_out/r2_08_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_08_bench.json')); print({k: round(d[k],3) for k in ('value','ms_per_step','loss')}, round(d['e2e']['value']), d['attn_roofline'])" 2>&1 | tail -1 | cut -c1-900
timeout -k 10 600 python tools/attn_sweep.py > gpurun_out/r2_08_attention_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2_08_attention_sweep.txt | cut -c1-150
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_08_trace.log 2>&1; tail -3 gpurun_out/r2_08_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_08_trace_summary.txt 2>&1; head -40 gpurun_out/r2_08_trace_summary.txt
