#!/usr/bin/env python
"""One attention-core shape launched a few times (ncu target).  usage: one_attn.py N T [fwd|bwd|row1f|row1b] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, ops  # noqa: E402

N, T = int(sys.argv[1]), int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "bwd"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
_lib.require_device()
BF = torch.bfloat16
H, d, C = 8, 64, 512
M = N * T
torch.manual_seed(0)
qkv = torch.randn(M, 3 * C, device="cuda").relu().to(BF)
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
graph = (torch.rand(N, T, T, device="cuda") < 0.3).float()
graph[:, torch.arange(T), torch.arange(T)] = 1
on = torch.ones(M, device="cuda")
dout = torch.randn(M, C, device="cuda")
dqkv = torch.empty(M, 3 * C, device="cuda", dtype=BF)
db = torch.zeros(3, C, device="cuda")
q1 = torch.randn(N, C, device="cuda").relu().to(BF)
g1 = torch.ones(N, 1, T, device="cuda")
on1 = torch.ones(N, device="cuda")
dout1 = torch.randn(N, C, device="cuda")
dq1 = torch.empty(N, C, device="cuda", dtype=BF)
# as the training step drives them: bit-packed graph, forward row statistics reused by the backward
bits = ops.pack_graph_bits(graph)
stats = torch.empty(H * N * T * 4, device="cuda")
if kind == "fwd":
    fn = lambda: ops.graph_attention_fwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, False, 0, graph_bits=bits, stats=stats)  # noqa: E731
elif kind == "bwd":
    fwd_out, _ = ops.graph_attention_fwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, False, 0, graph_bits=bits, stats=stats)
    fn = lambda: ops.graph_attention_bwd(q, k, v, graph, on, on, N, H, T, T, d, False, 1, dout, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:],  # noqa: E731
                                         dbq=db[0], dbk=db[1], dbv=db[2], graph_bits=bits, stats=stats, fwd_out=fwd_out)
elif kind == "row1f":
    fn = lambda: ops.graph_attention_fwd(q1, k, v, g1, on, on1, N, H, 1, T, d, False, 1, False, 1)  # noqa: E731
else:
    fn = lambda: ops.graph_attention_bwd(q1, k, v, g1, on, on1, N, H, 1, T, d, False, 1, dout1, dq1, dqkv[:, C:2 * C], dqkv[:, 2 * C:],  # noqa: E731
                                         dbq=db[0], dbk=db[1], dbv=db[2])
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
print(f"attn {kind} N={N} T={T}: us per launch {['%.1f' % t for t in ts]}")
