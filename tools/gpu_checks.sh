#!/bin/bash
# One gpurun call: every GPU test group in its own process (a trapped kernel must not poison the other groups),
# then smoke and a short bench.  Logs land in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.log
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400)" | tee -a gpurun_out/summary.log
}
rm -f gpurun_out/summary.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.log 2>&1
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 240"
run t_elem 600 $PT tests/test_gpu_kernels.py -k "masks or gather or casts or layernorm or loss"
run t_gemm_pair 600 $PT tests/test_gpu_kernels.py -k "pair_kernel"
run t_gemm_k 600 $PT tests/test_gpu_kernels.py -k "gemm_kmajor"
run t_gemm_epi 300 $PT tests/test_gpu_kernels.py -k "gemm_epilogues"
run t_gemm_wgrad 300 $PT tests/test_gpu_kernels.py -k "wgrad or dgrad"
run t_attn_simt 600 $PT tests/test_gpu_kernels.py -k "attention_forward and 1-"
run t_attn_tc 600 $PT tests/test_gpu_kernels.py -k "attention_forward and 0-"
run t_attn_bwd 600 $PT tests/test_gpu_kernels.py -k "attention_backward"
run t_parity 900 $PT tests/test_gpu_parity.py
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_eager 600 python bench.py --mode eager --steps 5 --warmup 3 --no-cpu-baseline
run bench_graph 600 python bench.py --mode graph --steps 10 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.log
