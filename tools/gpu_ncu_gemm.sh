#!/bin/bash
# ncu --set full capture of the CTA-pair GEMM on the dominant shape (FFN conv1 forward, symbolic branch).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ARGS="${1:-16384 2048 512 fwd 4}"
TAG="${2:-gemm2_ffn1}"
python tools/one_gemm.py $ARGS > gpurun_out/one_gemm_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 1 -c 2 -f -o gpurun_out/prof_${TAG} python tools/one_gemm.py $ARGS > gpurun_out/ncu_${TAG}.log 2>&1
echo "rc=$?"; cat gpurun_out/one_gemm_${TAG}.log; tail -3 gpurun_out/ncu_${TAG}.log
python tools/one_gemm.py 16384 2048 4096 fwd 4
python tools/one_gemm.py 16384 2048 512 fwd 4
python tools/one_gemm.py 16384 512 2048 res 4
python tools/one_gemm.py 65536 2048 512 fwd 4
