#!/bin/bash
# single GPU: cluster GEMM with 8 epilogue warps -- GPU suite, bench, decoder timings from the step timeline
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests > gpurun_out/r2_26_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_26_pytest.log)"
grep -E "^E  |FAILED" gpurun_out/r2_26_pytest.log | cut -c1-300 | head -12
timeout -k 10 200 python bench.py --no-extras --steps 40 > gpurun_out/r2_26_bench.json 2> gpurun_out/r2_26_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_26_bench.json')); print({k: round(d[k],3) for k in ('value','ms_per_step','loss')}, round(d['e2e']['value']))"
timeout -k 10 200 python tools/trace_step.py > gpurun_out/r2_26_trace.log 2>&1; tail -1 gpurun_out/r2_26_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_26_trace_summary.txt 2>&1; grep -E "span|rowln|decoder|loss " gpurun_out/r2_26_trace_summary.txt
