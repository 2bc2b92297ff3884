#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KIND="${1:-bwd}"; REGEX="${2:-attn_bwd_tc}"; TAG="${3:-attn_bwd}"
python tools/one_attn.py 128 128 $KIND 4 > gpurun_out/one_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 1 -c 1 -f -o gpurun_out/prof_${TAG} python tools/one_attn.py 128 128 $KIND 4 > gpurun_out/ncu_${TAG}.log 2>&1
echo "rc=$?"; cat gpurun_out/one_${TAG}.log; tail -2 gpurun_out/ncu_${TAG}.log
