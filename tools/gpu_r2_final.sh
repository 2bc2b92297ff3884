#!/bin/bash
# final single-GPU pass of the round: GPU tests, smoke, default bench line (all legs), reference arm, dropout line, sweep, timeline, ncu launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests > gpurun_out/r2_final_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_final_pytest.log)"
grep -E "^E  |FAILED" gpurun_out/r2_final_pytest.log | cut -c1-400 | head -20
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final_smoke.log | cut -c1-400
timeout -k 10 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_final_bench.json
python -c "
import json
d=json.load(open('gpurun_out/r2_final_bench.json'))
print('e2e', d['e2e'], 'step_roofline', round(d['step_roofline']['frac'],4), 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), 'launches', d['gpu_launches'])
print({k: (d[k] if not isinstance(d[k], dict) else {a: b for a, b in d[k].items() if a != 'note'}) for k in ('dense_tables','inference','stock_gpu_baseline','cpu_baseline','hbm_kernels') if k in d})"
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "reference arm rc=$?"; cut -c1-250 gpurun_out/r2_final_bench_reference.json
timeout -k 10 300 python bench.py --dropout 0.5 --no-extras > gpurun_out/r2_final_bench_dropout.json 2> gpurun_out/r2_final_bench_dropout.err; echo "dropout rc=$?"; cut -c1-250 gpurun_out/r2_final_bench_dropout.json
timeout -k 10 300 python tools/trace_step.py > gpurun_out/r2_final_trace.log 2>&1; tail -2 gpurun_out/r2_final_trace.log
python tools/summarize_trace.py gpurun_out/trace_step.json.gz > gpurun_out/r2_final_trace_summary.txt 2>&1; sed -n 1,14p gpurun_out/r2_final_trace_summary.txt
python bench.py --mode eager --steps 1 --warmup 3 --no-extras > gpurun_out/r2_final_plain_eager.log 2>&1 && \
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2_final.csv \
  python bench.py --mode eager --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_launches_r2_final.log 2>&1
echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r2_final.csv > gpurun_out/launches_r2_final.txt 2>&1; head -45 gpurun_out/launches_r2_final.txt
