#!/usr/bin/env python
"""The HBM-bound kernels of the path (and the dominant GEMM shape), launched a few times each for an `ncu --set full` capture of
their DRAM counters (north_star: "achieved HBM GB/s for LayerNorm and gather"): residual+LayerNorm forward / backward at the
symbolic branch's row count, the word-row gather, the bias-gradient column sums, the CTA-pair GEMM at conv1 M = 16384.
usage: python tools/hbm_kernels.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402

_lib.require_device()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
BF = torch.bfloat16
rows, C = 128 * 128, 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
x, res = torch.randn(rows, C, device="cuda"), torch.randn(rows, C, device="cuda")
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
dg, db, dxs = (torch.zeros(C, device="cuda") for _ in range(3))
dy = torch.randn(rows, C, device="cuda")
_, pre, _, _ = ops.layernorm_fwd(x, res, gamma, beta, 1e-8, True, False, False)
table = torch.randn(407000, 300, device="cuda")
idx = torch.randint(0, 407000, (1 << 18,), device="cuda")
dh = torch.randn(rows, 2048, device="cuda").to(BF)
cs = torch.zeros(2048, device="cuda")
a, w, bias = torch.randn(rows, C, device="cuda").to(BF), torch.randn(2048, C, device="cuda").to(BF), torch.randn(2048, device="cuda")
h = torch.empty(rows, 2048, device="cuda", dtype=BF)
for _ in range(reps):
    for fn in (lambda: ops.layernorm_fwd(x, res, gamma, beta, 1e-8, True, True, True),
               lambda: ops.layernorm_bwd(dy, pre, gamma, 1e-8, dg, db, want_bf16=True, dxsum=dxs),
               lambda: ops.gather_rows(table, idx, want_f32=True, want_bf16=True),
               lambda: ops.colsum_bf16(dh, cs),
               lambda: ops.gemm(a, w, rows, 2048, C, bias=bias, relu=True, out_bf16=h)):
        flush.zero_()
        fn()
torch.cuda.synchronize()
print("ok")
