#!/bin/bash
# step bench + ncu --set full of the attention forward / backward kernels as the step drives them (bit-packed graph, statistics)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-230
for T in 128 56; do
  for KIND in fwd bwd; do
    python tools/one_attn.py 128 $T $KIND 6 2>&1 | tail -1
    timeout -k 10 200 ncu --set full --clock-control none --import-source on -k regex:attn_${KIND}_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_${KIND}_T$T python tools/one_attn.py 128 $T $KIND 4 > gpurun_out/ncu_attn_${KIND}_T$T.log 2>&1
    echo "ncu $KIND T=$T rc=$?"
  done
done
