#!/usr/bin/env python
"""In-graph cost of one link of the decoder's kernel chain: N stream-ordered launches of the same small kernel captured in a
CUDA graph, replay time / N.  usage: chain_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402

_lib.require_device()
BF = torch.bfloat16
torch.manual_seed(0)
NL = 48


def chain(name, fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL):
                fn(i)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) * 1e3 / 5 / NL:7.2f} us per link", flush=True)


C = 512
x = torch.randn(128, 2048, device="cuda").to(BF)
# distinct weights per link (cold-ish, like the step: every layer has its own)
for name, N, K in (("dec qkv  M=128 N=1536 K=512", 1536, 512), ("dec q    M=128 N=512  K=512", 512, 512), ("dec ffn1 M=128 N=2048 K=512", 2048, 512),
                   ("dec ffn2 M=128 N=512  K=2048", 512, 2048)):
    ws = [torch.randn(N, K, device="cuda").to(BF) for _ in range(NL)]
    bias = torch.randn(N, device="cuda")
    out = torch.empty(128, N, device="cuda", dtype=BF)
    chain("gemm fwd " + name, lambda i: ops.gemm(x[:, :K], ws[i], 128, N, K, bias=bias, relu=True, out_bf16=out))
    dy = torch.randn(128, N, device="cuda").to(BF)
    dx = torch.empty(128, K, device="cuda")
    chain("gemm dgrad " + name, lambda i: ops.gemm(dy, ws[i], 128, K, N, b_mn=True, out_f32=dx))
    dws = [torch.zeros(N, K, device="cuda") for _ in range(NL)]
    chain("gemm wgrad " + name, lambda i: ops.wgrad(dy, x[:, :K].contiguous() if False else x[:, :K], N, K, dws[i]))
xf = torch.randn(128, C, device="cuda")
g_, b_ = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
chain("res+LN fwd rows=128", lambda i: ops.layernorm_fwd(xf, xf, g_, b_, 1e-8, True, True, True))
dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
chain("LN bwd rows=128", lambda i: ops.layernorm_bwd(xf, xf, g_, 1e-8, dg, db, want_bf16=True))
N, T, H, d = 128, 128, 8, 64
kv = torch.randn(N * T, 2 * C, device="cuda").relu().to(BF)
q1 = torch.randn(N, C, device="cuda").relu().to(BF)
g1 = torch.ones(N, 1, T, device="cuda")
on, on1 = torch.ones(N * T, device="cuda"), torch.ones(N, device="cuda")
chain("attn row1 fwd Tk=128", lambda i: ops.graph_attention_fwd(q1, kv[:, :C], kv[:, C:], g1, on, on1, N, H, 1, T, d, False, 1, False, 1))
dq1 = torch.empty(N, C, device="cuda", dtype=BF)
dkv = torch.empty(N * T, 2 * C, device="cuda", dtype=BF)
do1 = torch.randn(N, C, device="cuda")
chain("attn row1 bwd Tk=128", lambda i: ops.graph_attention_bwd(q1, kv[:, :C], kv[:, C:], g1, on, on1, N, H, 1, T, d, False, 1, do1, dq1, dkv[:, :C], dkv[:, C:]))
chain("attn row1 fwd Tk=1 (self)", lambda i: ops.graph_attention_fwd(q1, q1, q1, None, on1, on1, N, H, 1, 1, d, True, 0, False, 1))
