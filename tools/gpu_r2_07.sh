#!/bin/bash
# round 2, call 7: new attention shapes (tests + sweep), dropout-0.5 bench, then ncu: launch list of one eager step + --set full of the HBM kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 tests/test_gpu_kernels.py -k "small_heads or statistics" > gpurun_out/r2_07_attn_tests.log 2>&1; echo "attention tests rc=$? $(tail -n 1 gpurun_out/r2_07_attn_tests.log)"
grep -E "^E  |FAILED" gpurun_out/r2_07_attn_tests.log | cut -c1-300 | head
timeout -k 10 600 python tools/attn_sweep.py > gpurun_out/r2_07_attention_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2_07_attention_sweep.txt | cut -c1-150
timeout -k 10 300 python bench.py --no-extras --dropout 0.5 > gpurun_out/r2_07_bench_dropout.json 2> gpurun_out/r2_07_bench_dropout.err; echo "dropout bench rc=$?"; cut -c1-200 gpurun_out/r2_07_bench_dropout.json; tail -3 gpurun_out/r2_07_bench_dropout.err
bash tools/gpu_r2_ncu.sh
