#!/usr/bin/env python
"""One GEMM shape launched a few times (ncu target).  usage: one_gemm.py M N K [kind=fwd|res|dgrad_gate|wgrad] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
kind = sys.argv[4] if len(sys.argv) > 4 else "fwd"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
_lib.require_device()
BF = torch.bfloat16
torch.manual_seed(0)
if kind == "fwd":
    a, w, bias = torch.randn(M, K, device="cuda").to(BF), torch.randn(N, K, device="cuda").to(BF), torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=BF)
    fn = lambda: ops.gemm(a, w, M, N, K, bias=bias, relu=True, out_bf16=out)  # noqa: E731
elif kind == "res":
    a, w, bias = torch.randn(M, K, device="cuda").to(BF), torch.randn(N, K, device="cuda").to(BF), torch.randn(N, device="cuda")
    res, out = torch.randn(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    fn = lambda: ops.gemm(a, w, M, N, K, bias=bias, res=res, out_f32=out)  # noqa: E731
elif kind == "dgrad_gate":
    dy, w = torch.randn(M, K, device="cuda").to(BF), torch.randn(K, N, device="cuda").to(BF)
    gate, out, cs = torch.randn(M, N, device="cuda").to(BF), torch.empty(M, N, device="cuda", dtype=BF), torch.zeros(N, device="cuda")
    fn = lambda: ops.gemm(dy, w, M, N, K, b_mn=True, gate=gate, out_bf16=out, colsum=cs)  # noqa: E731
else:
    dy, x, out = torch.randn(K, M, device="cuda").to(BF), torch.randn(K, N, device="cuda").to(BF), torch.zeros(M, N, device="cuda")
    fn = lambda: ops.wgrad(dy, x, M, N, out)  # noqa: E731
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"{kind} M={M} N={N} K={K}: us per launch {['%.1f' % t for t in ts]}  best {2.0 * M * N * K / min(ts) / 1e6:.1f} TFLOP/s")
