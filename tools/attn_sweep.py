#!/usr/bin/env python
"""BASELINE.json configs[4]: graph-masked attention microbenchmark sweep -- nodes 36..256, heads 8..16, d_model 512..1024, dense vs
sparse adjacency -- forward and backward of the fused attention core (bit-packed graph, forward statistics), each launch timed
alone with CUDA events after an L2 flush.  Prints time, dense-equivalent TFLOP/s and its share of the measured bf16 peak, and the
algorithmic GB/s and its share of the measured HBM peak.  usage: python tools/attn_sweep.py [batch=128]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402
from savqa_b200.functional import tc_attention_fits  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
_lib.require_device()
BF = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.manual_seed(0)


def timeit(fn, iters=8):
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


print(f"batch {N} samples; peaks: {peaks['bf16_tflops']} TFLOP/s bf16 (burst), {peaks['hbm_gbs']} GB/s HBM")
print(f"{'d_model':>7s} {'heads':>5s} {'d':>4s} {'T':>4s} {'graph':>7s} | {'fwd us':>7s} {'TF/s':>7s} {'%tensor':>8s} {'%HBM':>6s} | {'bwd us':>7s} {'TF/s':>7s} {'%tensor':>8s} {'%HBM':>6s}")
for C, H in ((512, 8), (512, 16), (1024, 16), (1024, 8)):
    d = C // H
    for T in (36, 56, 100, 128, 256):
        if not tc_attention_fits(d, T):
            continue
        M = N * T
        qkv = torch.randn(M, 3 * C, device="cuda").relu().to(BF)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        on = torch.ones(M, device="cuda")
        dout = torch.randn(M, C, device="cuda")
        dqkv = torch.empty(M, 3 * C, device="cuda", dtype=BF)
        for label, dens in (("dense", 1.0), ("10%", 0.1)):
            graph = (torch.rand(N, T, T, device="cuda") < dens).float()
            graph[:, torch.arange(T), torch.arange(T)] = 1
            bits = ops.pack_graph_bits(graph)
            stats = torch.empty(H * N * T * 4, device="cuda")
            fwd = lambda: ops.graph_attention_fwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, False, 0, graph_bits=bits, stats=stats)  # noqa: E731
            o, _ = fwd()
            us_f = timeit(fwd)
            ff, fb = 4.0 * N * H * T * T * d, M * 3 * C * 2 + N * T * T / 8 + M * C * 4
            row = f"{C:7d} {H:5d} {d:4d} {T:4d} {label:>7s} | {us_f:7.1f} {ff / us_f / 1e6:7.1f} {100 * ff / us_f / 1e6 / peaks['bf16_tflops']:7.2f}% {100 * fb / us_f / 1e3 / peaks['hbm_gbs']:5.1f}% |"
            if ops.tc_attention_bwd_fits(d, T, T):
                bwd = lambda: ops.graph_attention_bwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, dout, dqkv[:, :C], dqkv[:, C:2 * C],  # noqa: E731
                                                      dqkv[:, 2 * C:], graph_bits=bits, stats=stats, fwd_out=o)
                us_b = timeit(bwd)
                bf_, bb = 10.0 * N * H * T * T * d, M * 3 * C * 2 * 2 + N * T * T / 8 + 2 * M * C * 4
                row += f" {us_b:7.1f} {bf_ / us_b / 1e6:7.1f} {100 * bf_ / us_b / 1e6 / peaks['bf16_tflops']:7.2f}% {100 * bb / us_b / 1e3 / peaks['hbm_gbs']:5.1f}%"
                if not ops._bwd_one_cta_fits(64 if d == 32 else d, T, T):
                    row += "  (tiled: query x key tiles of <= 128 on the forward statistics)"
            else:
                row += "   (backward: CUDA-core engine for this shape)"
            print(row, flush=True)
