#!/usr/bin/env python
"""The step's big GEMM shapes timed two ways: cold (L2 flushed before each launch) and hot (20 launches back to back / 20).
usage: [SAVQA_LIB=...] python tools/gemm_variants.py [tag]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402

BF = torch.bfloat16
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(_lib.LIB_PATH)
_lib.require_device()
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def cold(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def hot(fn, n=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


C, Hd = 512, 2048
for M in (7168, 16384):
    x = torch.randn(M, Hd, device="cuda").to(BF)
    for name, N, K in (("qkv fwd", 3 * C, C), ("ffn1 fwd", Hd, C), ("ffn2 fwd", C, Hd)):
        w = torch.randn(N, K, device="cuda").to(BF)
        bias = torch.randn(N, device="cuda")
        if N == C:
            res, out = torch.randn(M, C, device="cuda"), torch.empty(M, N, device="cuda")
            fn = lambda: ops.gemm(x[:, :K], w, M, N, K, bias=bias, res=res, out_f32=out)  # noqa: E731
        else:
            out = torch.empty(M, N, device="cuda", dtype=BF)
            fn = lambda: ops.gemm(x[:, :K], w, M, N, K, bias=bias, relu=True, out_bf16=out)  # noqa: E731
        c, h = cold(fn), hot(fn)
        fl = 2.0 * M * N * K
        print(f"[{tag}] {name:10s} M={M:5d} N={N:4d} K={K:4d}  cold {c:6.1f} us {fl / c / 1e6:7.1f} TF/s   hot {h:6.1f} us {fl / h / 1e6:7.1f} TF/s", flush=True)
    # ffn2 dgrad (gate + colsum, bf16 out, K=512) and qkv dgrad (res, f32 out)
    w = torch.randn(C, Hd, device="cuda").to(BF)
    dy = torch.randn(M, C, device="cuda").to(BF)
    gate, out, cs = torch.randn(M, Hd, device="cuda").to(BF), torch.empty(M, Hd, device="cuda", dtype=BF), torch.zeros(Hd, device="cuda")
    fn = lambda: ops.gemm(dy, w, M, Hd, C, b_mn=True, gate=gate, out_bf16=out, colsum=cs)  # noqa: E731
    c, h = cold(fn), hot(fn)
    fl = 2.0 * M * Hd * C
    print(f"[{tag}] {'ffn2 dgrad':10s} M={M:5d} N={Hd:4d} K={C:4d}  cold {c:6.1f} us {fl / c / 1e6:7.1f} TF/s   hot {h:6.1f} us {fl / h / 1e6:7.1f} TF/s", flush=True)
    dyw = torch.randn(M, Hd, device="cuda").to(BF)
    xx = torch.randn(M, C, device="cuda").to(BF)
    o = torch.zeros(Hd, C, device="cuda")
    fn = lambda: ops.wgrad(dyw, xx, Hd, C, o)  # noqa: E731
    c, h = cold(fn), hot(fn)
    print(f"[{tag}] {'ffn1 wgrad':10s} M={Hd:5d} N={C:4d} K={M:4d}  cold {c:6.1f} us {fl / c / 1e6:7.1f} TF/s   hot {h:6.1f} us {fl / h / 1e6:7.1f} TF/s", flush=True)
