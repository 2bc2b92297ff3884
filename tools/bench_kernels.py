#!/usr/bin/env python
"""Per-kernel microbenchmarks of the step's shapes on one B200 (CUDA events, L2 flushed between iterations).
Prints one line per case: time, achieved TFLOP/s or GB/s and the fraction of the measured peak (MEASURED_PEAKS.json).
usage: python tools/bench_kernels.py [gemm] [attn] [ln] [adam]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402

BF, F32 = torch.bfloat16, torch.float32
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush_buf = None


def timeit(fn, iters=20, warmup=3):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()  # 256 MB > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, us, flops=None, bytes_=None):
    s = f"{name:58s} {us:9.1f} us"
    if flops:
        tf = flops / us / 1e6
        s += f"  {tf:8.1f} TFLOP/s ({100 * tf / peaks['bf16_tflops']:5.1f}% of measured burst)"
    if bytes_:
        gb = bytes_ / us / 1e3
        s += f"  {gb:8.1f} GB/s ({100 * gb / peaks['hbm_gbs']:5.1f}% of measured HBM)"
    print(s, flush=True)


def bench_gemm():
    C, Hd = 512, 2048
    for M in (7168, 16384):
        x = torch.randn(M, Hd, device="cuda").to(BF)
        for name, N, K in (("qkv fwd", 3 * C, C), ("kv fwd", 2 * C, C), ("ffn1 fwd", Hd, C), ("ffn2 fwd", C, Hd)):
            w = torch.randn(N, K, device="cuda").to(BF)
            bias = torch.randn(N, device="cuda")
            if N == C:
                res = torch.randn(M, C, device="cuda")
                out = torch.empty(M, N, device="cuda")
                us = timeit(lambda: ops.gemm(x[:, :K], w, M, N, K, bias=bias, res=res, out_f32=out))
                by = M * K * 2 + N * K * 2 + 2 * M * N * 4
            else:
                out = torch.empty(M, N, device="cuda", dtype=BF)
                us = timeit(lambda: ops.gemm(x[:, :K], w, M, N, K, bias=bias, relu=True, out_bf16=out))
                by = M * K * 2 + N * K * 2 + M * N * 2
            report(f"gemm {name:10s} M={M} N={N} K={K}", us, 2.0 * M * N * K, by)
        # dgrad: dX[M,K] = dY[M,N] W[N,K]  (B MN-major)
        for name, N, K, kind in (("qkv dgrad", 3 * C, C, "res"), ("ffn1 dgrad", Hd, C, "res"), ("ffn2 dgrad", C, Hd, "gate")):
            w = torch.randn(N, K, device="cuda").to(BF)
            dy = torch.randn(M, N, device="cuda").to(BF)
            if kind == "res":
                res = torch.randn(M, K, device="cuda")
                out = torch.empty(M, K, device="cuda")
                us = timeit(lambda: ops.gemm(dy, w, M, K, N, b_mn=True, res=res, out_f32=out))
                by = M * N * 2 + N * K * 2 + 2 * M * K * 4
            else:
                gate = torch.randn(M, K, device="cuda").to(BF)
                out = torch.empty(M, K, device="cuda", dtype=BF)
                cs = torch.zeros(K, device="cuda")
                us = timeit(lambda: ops.gemm(dy, w, M, K, N, b_mn=True, gate=gate, out_bf16=out, colsum=cs))
                by = M * N * 2 + N * K * 2 + 2 * M * K * 2
            report(f"gemm {name:10s} M={M} N={K} K={N}", us, 2.0 * M * N * K, by)
        for name, N, K in (("qkv wgrad", 3 * C, C), ("ffn1 wgrad", Hd, C), ("ffn2 wgrad", C, Hd)):
            dy = torch.randn(M, N, device="cuda").to(BF)
            xx = torch.randn(M, K, device="cuda").to(BF)
            out = torch.zeros(N, K, device="cuda")
            us = timeit(lambda: ops.wgrad(dy, xx, N, K, out))
            report(f"gemm {name:10s} M={N} N={K} K={M}", us, 2.0 * M * N * K, M * (N + K) * 2 + N * K * 4)
    # decoder-sized and head GEMMs
    for name, M, N, K in (("dec qkv", 128, 1536, 512), ("dec ffn1", 128, 2048, 512), ("dec ffn2", 128, 512, 2048), ("head", 128, 1845, 512)):
        x = torch.randn(M, K, device="cuda").to(BF)
        w = torch.randn(N, K, device="cuda").to(BF)
        out = torch.empty(M, N, device="cuda")
        us = timeit(lambda: ops.gemm(x, w, M, N, K, out_f32=out))
        report(f"gemm {name:10s} M={M} N={N} K={K}", us, 2.0 * M * N * K)


def bench_attn():
    H, d, C = 8, 64, 512
    for N, T in ((128, 56), (128, 128), (256, 120)):
        M = N * T
        qkv = torch.randn(M, 3 * C, device="cuda").relu().to(BF)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        graph = (torch.rand(N, T, T, device="cuda") < 0.3).float()
        graph[:, torch.arange(T), torch.arange(T)] = 1
        on = torch.ones(N * T, device="cuda")
        bits = ops.pack_graph_bits(graph)
        us = timeit(lambda: ops.graph_attention_fwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, False, 0, graph_bits=bits))
        by = M * 3 * C * 2 + N * T * T * 4 + M * C * 4
        report(f"attn fwd tc   N={N} T={T}", us, 4.0 * N * H * T * T * d, by)
        dout = torch.randn(M, C, device="cuda")
        dqkv = torch.empty(M, 3 * C, device="cuda", dtype=BF)
        db = torch.zeros(3, C, device="cuda")
        us = timeit(lambda: ops.graph_attention_bwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, dout, dqkv[:, :C], dqkv[:, C:2 * C],
                                                    dqkv[:, 2 * C:], dbq=db[0], dbk=db[1], dbv=db[2], graph_bits=bits))
        by = M * 3 * C * 2 * 2 + N * T * T * 4 + M * C * 4
        report(f"attn bwd tc   N={N} T={T}", us, 10.0 * N * H * T * T * d, by)
        # decoder cross-attention: one query per sample
        q1 = torch.randn(N, C, device="cuda").relu().to(BF)
        g1 = torch.ones(N, 1, T, device="cuda")
        on1 = torch.ones(N, device="cuda")
        us = timeit(lambda: ops.graph_attention_fwd(q1, k, v, g1, on, on1, N, H, 1, T, d, False, 1, False, 1))
        report(f"attn fwd row1 N={N} Tk={T}", us, None, M * 2 * C * 2)
        dout1 = torch.randn(N, C, device="cuda")
        dq1 = torch.empty(N, C, device="cuda", dtype=BF)
        dkv = torch.empty(M, 2 * C, device="cuda", dtype=BF)
        us = timeit(lambda: ops.graph_attention_bwd(q1, k, v, g1, on, on1, N, H, 1, T, d, False, 1, dout1, dq1, dkv[:, :C], dkv[:, C:],
                                                    dbq=db[0], dbk=db[1], dbv=db[2]))
        report(f"attn bwd row1 N={N} Tk={T}", us, None, M * 2 * C * 2 * 2)


def bench_ln():
    C = 512
    for rows in (7168, 16384):
        x = torch.randn(rows, C, device="cuda")
        r = torch.randn(rows, C, device="cuda")
        g = torch.rand(C, device="cuda") + 0.5
        b = torch.randn(C, device="cuda")
        us = timeit(lambda: ops.layernorm_fwd(x, r, g, b, 1e-8, True, True, True))
        report(f"res+LN fwd rows={rows} (pre, y, y_bf16)", us, None, rows * C * (4 * 4 + 2))
        dg, db, ds = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        us = timeit(lambda: ops.layernorm_bwd(x, r, g, 1e-8, dg, db, want_bf16=True, dxsum=ds))
        report(f"LN bwd rows={rows} (dx, dx_bf16)", us, None, rows * C * (3 * 4 + 2))


def bench_adam():
    n = 88_900_000 // 8 * 8
    p, g, m, v = (torch.zeros(n, device="cuda") for _ in range(4))
    mirror = torch.zeros(n, device="cuda", dtype=BF)
    us = timeit(lambda: ops.adam_step(p, g, m, v, 1e-4, 0.9, 0.999, 1e-8, 1, param_bf16=mirror), iters=5)
    report(f"adam n={n} (+bf16 mirror)", us, None, n * 30)


if __name__ == "__main__":
    _lib.require_device()
    which = sys.argv[1:] or ["gemm", "attn", "ln", "adam"]
    torch.manual_seed(0)
    for w in which:
        {"gemm": bench_gemm, "attn": bench_attn, "ln": bench_ln, "adam": bench_adam}[w]()
