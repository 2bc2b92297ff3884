#!/bin/bash
# After gpu_checks.sh: ncu launch list of one eager step (cold-cache, serialised per-launch times).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-rX}
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --mode eager --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.txt 2>&1
head -50 gpurun_out/launches_${TAG}.txt
