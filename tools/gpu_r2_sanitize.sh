#!/bin/bash
# usage: tools/gpu_r2_sanitize.sh memcheck|racecheck|synccheck   (one compute-sanitizer tool per gpurun call)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOL=${1:-memcheck}
python tools/sanitize_step.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_plain.log | cut -c1-300
timeout -k 10 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_step.py > gpurun_out/sanitize_${TOOL}_r2.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_${TOOL}_r2.log | head -20
tail -3 gpurun_out/sanitize_${TOOL}_r2.log | cut -c1-300
