#!/bin/bash
# ncu call: (1) launch list of one eager step, (2) --set full of the HBM-bound kernels + the dominant GEMM shape
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --mode eager --steps 1 --warmup 3 --no-extras > gpurun_out/r2_plain_eager.log 2>&1 && \
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --mode eager --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_launches_r2.log 2>&1
echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r2.csv > gpurun_out/launches_r2.txt 2>&1; head -60 gpurun_out/launches_r2.txt
python tools/hbm_kernels.py 3 > gpurun_out/hbm_plain.log 2>&1 && \
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'ln_fwd|ln_bwd|gather_rows|colsum|gemm2_bf16' -c 15 -f -o gpurun_out/prof_hbm_r2 \
  python tools/hbm_kernels.py 3 > gpurun_out/ncu_hbm_r2.log 2>&1
echo "ncu hbm rc=$?"; tail -3 gpurun_out/ncu_hbm_r2.log
