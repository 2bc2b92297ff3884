#!/bin/bash
# single GPU: static SM split between the two branch streams, a few points around the default (56,92)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 56,92 60,88 52,96 64,84; do
  SAVQA_BRANCH_SMS=$v timeout -k 10 200 python bench.py --no-extras --steps 40 > gpurun_out/r2_split_$v.json 2> gpurun_out/r2_split_$v.err; echo -n "split $v rc=$? "
  python -c "
import json
d=json.load(open('gpurun_out/r2_split_$v.json')); print({k: round(d[k],3) for k in ('value','ms_per_step')}, round(d['e2e']['value']))"
done
