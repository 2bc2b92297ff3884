#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests/test_gpu_parity.py > gpurun_out/t_par.log 2>&1; echo "parity rc=$? $(tail -n 1 gpurun_out/t_par.log)"
grep -E "^E  |FAILED" gpurun_out/t_par.log | head
for off in 0 1; do
  if [ $off = 1 ]; then export SAVQA_ROW1_SPLIT_OFF=1; fi
  echo "== SPLIT_OFF=$off"
  for T in 128 56; do for KIND in row1f row1b; do python tools/one_attn.py 128 $T $KIND 8 2>&1 | tail -1; done; done
  timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
