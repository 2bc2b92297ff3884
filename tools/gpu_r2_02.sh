#!/bin/bash
# round 2, call 2: GPU tests (MIL_NCE, compact hand-off, full step, deferred Adam), smoke, default bench (whole step, compact inputs), round-1 workload for comparison
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 -x tests -s > gpurun_out/r2_02_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_02_pytest.log)"
grep -E "^E  |FAILED|engines:|headline step parity|^56 |^128 |^mil_nce" gpurun_out/r2_02_pytest.log | cut -c1-400 | head -30
timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|Error|error" | cut -c1-600
timeout -k 10 900 python bench.py > gpurun_out/r2_02_bench.json 2> gpurun_out/r2_02_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_02_bench.json; tail -5 gpurun_out/r2_02_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_02_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "loss", "gpu_launches")}, d["e2e"], d["step_roofline"]["frac"], d["roofline"]["frac"])
    for k in ("dense_tables", "inference", "stock_gpu_baseline", "cpu_baseline", "hbm_kernels", "attn_roofline"):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print("no bench line", e)
PY
timeout -k 10 300 python bench.py --step encoder --no-extras > gpurun_out/r2_02_bench_encoder.json 2>/dev/null; cut -c1-200 gpurun_out/r2_02_bench_encoder.json
