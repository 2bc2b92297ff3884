#!/bin/bash
# round 2, call 3: cluster row-LN GEMM kernel tests, fused decoder, full GPU suite, bench with the fused decoder on / off
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests/test_gpu_kernels.py -k "rowln or deferred" > gpurun_out/r2_03_rowln.log 2>&1; echo "rowln tests rc=$? $(tail -n 1 gpurun_out/r2_03_rowln.log)"
grep -E "^E  |FAILED|Error|error:" gpurun_out/r2_03_rowln.log | cut -c1-300 | head -20
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests -s > gpurun_out/r2_03_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_03_pytest.log)"
grep -E "^E  |FAILED|engines:|fused decoder vs chain|headline step parity" gpurun_out/r2_03_pytest.log | cut -c1-400 | head -30
timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|Error|error" | cut -c1-400
for F in 1 0; do
  SAVQA_FUSED_DECODER=$F timeout -k 10 300 python bench.py --no-extras > gpurun_out/r2_03_bench_fused$F.json 2> gpurun_out/r2_03_bench_fused$F.err; echo "bench fused=$F rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2_03_bench_fused$F.json')); print({k: d[k] for k in ('value','ms_per_step','loss','gpu_launches')}, d['e2e']['value'])" 2>&1 | tail -1
done
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_03_trace.log 2>&1; tail -12 gpurun_out/r2_03_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_03_trace_summary.txt 2>&1; head -50 gpurun_out/r2_03_trace_summary.txt
