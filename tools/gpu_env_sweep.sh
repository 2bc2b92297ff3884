#!/bin/bash
# usage: tools/gpu_env_sweep.sh "VAR=val VAR2=val" "VAR=val" ...   -- graph-mode step bench under each environment
set -u
cd "$(dirname "$0")/.."
for envs in "$@"; do
  echo "== $envs"
  env $envs timeout -k 10 300 python bench.py --mode graph --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
