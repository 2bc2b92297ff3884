#!/bin/bash
# Builds experiment variants of the library: tools/build_variants.sh NAME "<extra nvcc flags>" file.cu [file.cu ...]
# -> structured-alignment-vqa_b200/lib/variants/libsavqa_NAME.so (the listed sources recompiled with the flags, everything else
# linked from the regular build).  Select at run time with SAVQA_LIB=<path>.  Bring-up aid only; nothing ships from here.
set -eu
cd "$(dirname "$0")/../structured-alignment-vqa_b200"
name=$1; flags=$2; shift 2
mkdir -p lib/variants
objs=""
for f in api elementwise layernorm gemm_tcgen05 gemm2_tcgen05 attn_simt attn_tcgen05 attn_bwd_tcgen05; do
  use=lib/$f.o
  for s in "$@"; do
    if [ "$s" = "$f.cu" ]; then
      use=lib/variants/${f}_$name.o
      nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c csrc/$f.cu -o $use
    fi
  done
  objs="$objs $use"
done
nvcc -shared -o lib/variants/libsavqa_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcudart
echo lib/variants/libsavqa_$name.so
