#!/bin/bash
# full GPU test suite + attention microbench + graph-mode step bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/t_all.log)"
grep -E "^E  |FAILED" gpurun_out/t_all.log | head -20
for T in 128 56; do for KIND in fwd bwd row1f row1b; do python tools/one_attn.py 128 $T $KIND 6 2>&1 | tail -1; done; done
timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-230
