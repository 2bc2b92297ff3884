#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests/test_gpu_kernels.py -k "attention or masks or mask" > gpurun_out/t_k.log 2>&1; echo "kernel tests rc=$? $(tail -n 1 gpurun_out/t_k.log)"
grep -E "^E  |FAILED" gpurun_out/t_k.log | head
timeout -k 10 300 python tools/trace_step.py 2>&1 | tail -8
