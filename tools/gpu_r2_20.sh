#!/bin/bash
# single GPU: standalone GEMM roofline after the SM-limit fix, then ncu --set full of the decoder's cluster GEMM (LayerNorm epilogues)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --no-extras > gpurun_out/r2_20_roofline.json 2> gpurun_out/r2_20_roofline.err; echo "roofline rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_20_roofline.json')); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['shapes'])"
timeout -k 10 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "rowln_layernorm_modes" > gpurun_out/r2_20_rowln_plain.log 2>&1; echo "plain rc=$? $(tail -n 1 gpurun_out/r2_20_rowln_plain.log)"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rowln_gemm -c 8 -f -o gpurun_out/prof_rowln_r2 \
  python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "rowln_layernorm_modes" > gpurun_out/r2_20_ncu_rowln.log 2>&1
echo "ncu rowln rc=$?"; tail -3 gpurun_out/r2_20_ncu_rowln.log
