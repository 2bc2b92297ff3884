#!/bin/bash
# single-GPU validation of HEAD: GPU tests, smoke, default bench line (all legs), reference arm, stock-gpu arm, dropout line, sweep, trace
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests > gpurun_out/r2_12_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_12_pytest.log)"
grep -E "^E  |FAILED" gpurun_out/r2_12_pytest.log | cut -c1-400 | head -20
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_12_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_12_smoke.log | cut -c1-400
timeout -k 10 900 python bench.py > gpurun_out/r2_12_bench.json 2> gpurun_out/r2_12_bench.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r2_12_bench.json
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_12_bench_reference.json 2> gpurun_out/r2_12_bench_reference.err; echo "reference arm rc=$?"; cut -c1-700 gpurun_out/r2_12_bench_reference.json
timeout -k 10 300 python bench.py --dropout 0.5 --no-extras > gpurun_out/r2_12_bench_dropout.json 2> gpurun_out/r2_12_bench_dropout.err; echo "dropout rc=$?"; cut -c1-300 gpurun_out/r2_12_bench_dropout.json
timeout -k 10 600 python tools/attn_sweep.py > gpurun_out/r2_12_attention_sweep.txt 2>&1; echo "sweep rc=$?"; cut -c1-150 gpurun_out/r2_12_attention_sweep.txt
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_12_trace.log 2>&1; tail -3 gpurun_out/r2_12_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_12_trace_summary.txt 2>&1; head -50 gpurun_out/r2_12_trace_summary.txt
