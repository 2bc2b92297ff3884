#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, the default bench line + the reference arm, then (each only after its command
# exited 0 without ncu) the ncu launch list of one eager step, ncu --set full of the dominant GEMM and of the attention kernels,
# and the kernel timeline of the replayed graph.  usage: tools/gpu_final.sh TAG
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1_07}
timeout -k 10 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -x tests > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$? $(tail -n 1 gpurun_out/t_all.log)"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -n 1 gpurun_out/smoke.log | cut -c1-160)"
( time timeout -k 10 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
( time timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-160 gpurun_out/bench_reference.json
timeout -k 10 300 python tools/trace_step.py 2>&1 | tail -4
timeout -k 10 300 python tools/bench_kernels.py gemm attn ln adam > gpurun_out/kbench_${TAG}.log 2>&1; echo "kbench rc=$?"
# (round 1: with -c 6000 and 3 warm-up steps this pass ran into a 900 s limit -- ~15 GPU-minutes -- after the step had long been
#  written; one eager step is ~650 launches: stop ncu after the first few steps and cap the pass at 5 minutes)
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --mode eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu launch list rc=$?"
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.txt 2>&1; head -30 gpurun_out/launches_${TAG}.txt
python tools/one_gemm.py 16384 2048 512 fwd 4 > gpurun_out/one_gemm_${TAG}.log 2>&1 && \
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 1 -c 2 -f -o gpurun_out/prof_gemm2_conv1_${TAG} python tools/one_gemm.py 16384 2048 512 fwd 4 > gpurun_out/ncu_gemm_${TAG}.log 2>&1; echo "ncu gemm rc=$?"
for KIND in fwd bwd; do
  python tools/one_attn.py 128 128 $KIND 4 > gpurun_out/one_attn_${KIND}_${TAG}.log 2>&1 && \
  timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_${KIND}_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_${KIND}_${TAG} python tools/one_attn.py 128 128 $KIND 4 > gpurun_out/ncu_attn_${KIND}_${TAG}.log 2>&1; echo "ncu attn $KIND rc=$?"
done
