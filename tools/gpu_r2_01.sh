#!/bin/bash
# round 2, call 1: full GPU test suite (new full-size parity tests, deferred Adam), default bench line, reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -k 10 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 -x tests -s > gpurun_out/r2_01_pytest.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/r2_01_pytest.log)"
grep -E "^E  |FAILED|engines:|headline step parity|^56 |^128 " gpurun_out/r2_01_pytest.log | head -30
timeout -k 10 900 python bench.py > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_01_bench.json; tail -5 gpurun_out/r2_01_bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_01_reference.json 2>/dev/null; echo "reference rc=$?"; cut -c1-200 gpurun_out/r2_01_reference.json
