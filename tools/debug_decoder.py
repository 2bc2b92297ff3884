#!/usr/bin/env python
"""Fused decoder (functional.DecoderFn) vs the per-module chain on the GPU, intermediate by intermediate (debug aid)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch
from savqa_b200 import functional as Fn, ops, synthetic, train
cfg = dict(synthetic.GQA_SHAPED, ncls=256)
model = synthetic.build_model(cfg, vocab_rows=4000).cuda()
batch = {k: v.cuda() for k, v in synthetic.make_batch(cfg, 128, seed=13, vocab_rows=4000).items()}
rec = {}
orig = ops.gemm_rowln
def spy_rowln(a, b, M, N, K, mode, **kw):
    orig(a, b, M, N, K, mode, **kw)
    tag = f"rowln{len([k for k in rec if k.startswith('rowln')]):03d}_m{mode}_N{N}_K{K}"
    rec[tag] = {k: v.detach().float().clone() for k, v in kw.items() if k in ("y", "y_bf16", "dxg_bf16", "pre", "stats") and v is not None}
def run(fused):
    Fn.FUSED_DECODER = fused
    m = copy.deepcopy(model)
    tr = train.EncoderTrainer(m, lr=1e-4)
    tr.prepare(batch)
    tr.flat_grad.zero_()
    loss = tr._forward_backward(batch)
    Fn.join_wgrad_streams(); torch.cuda.synchronize()
    return tr, float(loss)
tr1, l1 = run(True)
tr0, l0 = run(False)
names = {id(p): k for k, p in tr1.model.named_parameters()}
names0 = {k: p for k, p in tr0.model.named_parameters()}
rows = []
for p in tr1.dense:
    k = names[id(p)]
    a, b = p.grad.float(), names0[k].grad.float()
    rows.append((float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max() / (b.abs().max() + 1e-12)), k))
rows.sort(reverse=True)
print("loss", l1, l0)
for r in rows[:40]:
    print(f"{r[0]:.3e} {r[1]:.3e} {r[2]}")
# repeat the fused run twice: run-to-run noise of the fused path itself
tr2, l2 = run(True)
rows = []
n2 = {k: p for k, p in tr2.model.named_parameters()}
for p in tr1.dense:
    k = names[id(p)]
    a, b = p.grad.float(), n2[k].grad.float()
    rows.append((float((a - b).norm() / (b.norm() + 1e-30)), k))
rows.sort(reverse=True)
print("fused vs fused (run-to-run):", rows[:5])
tr3, l3 = run(False)
rows = []
n3 = {k: p for k, p in tr3.model.named_parameters()}
for k, p in names0.items():
    if p.grad is not None and n3[k].grad is not None:
        rows.append((float((p.grad.float() - n3[k].grad.float()).norm() / (p.grad.float().norm() + 1e-30)), k))
rows.sort(reverse=True)
print("chain vs chain (run-to-run):", rows[:5])
