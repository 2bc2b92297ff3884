#!/bin/bash
# usage: tools/gpu_variants.sh name [name ...]   -- times each lib/variants/libsavqa_<name>.so (and "base" = the regular build)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=structured-alignment-vqa_b200/lib/variants
for name in "$@"; do
  if [ "$name" = base ]; then unset SAVQA_LIB; else export SAVQA_LIB=$PWD/$V/libsavqa_$name.so; fi
  case "$name" in
    diag*) ;;
    *) timeout -k 10 300 python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 -x tests/test_gpu_kernels.py -k "gemm" > gpurun_out/v_$name.test.log 2>&1; echo "[$name] gemm tests rc=$? $(tail -n 1 gpurun_out/v_$name.test.log)";;
  esac
  timeout -k 10 200 python tools/gemm_variants.py $name 2>&1 | tee gpurun_out/v_$name.log | grep "^\["
done
