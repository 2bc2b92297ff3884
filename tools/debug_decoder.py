#!/usr/bin/env python
"""Fused decoder (functional.DecoderFn) vs the per-module chain on the GPU: forward outputs layer by layer (debug aid)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch
from savqa_b200 import functional as Fn, ops, synthetic, train
cfg = dict(synthetic.GQA_SHAPED, ncls=256)
model = synthetic.build_model(cfg, vocab_rows=4000).cuda().eval()
batch = {k: v.cuda() for k, v in synthetic.make_batch(cfg, 128, seed=13, vocab_rows=4000).items()}
rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))
br = model.att_vis_grid
taps = {}
hooks = []
for i in range(6):
    for nm in ("dec_self_attention_%d", "dec_vanilla_attention_%d", "dec_feed_forward_%d"):
        m = getattr(br, nm % i)
        hooks.append(m.register_forward_hook(lambda mod, inp, out, key=nm % i: taps.__setitem__(key, out.detach().clone())))
orig = ops.gemm_rowln
spy = []
def spy_rowln(a, b, M, N, K, mode, **kw):
    orig(a, b, M, N, K, mode, **kw)
    spy.append((mode, N, K, {k: v.detach().clone() for k, v in kw.items() if k in ("y", "y_bf16", "pre", "act_bf16", "on", "stats") and v is not None}))
with torch.no_grad():
    Fn.FUSED_DECODER = False
    f0 = br(batch["vis_fea"], batch["vis_fea_mask"], batch["q_ipt"], batch["q_ipt_graph"], batch["q_ipt_mask"], True)
    Fn.FUSED_DECODER = True
    ops.gemm_rowln = spy_rowln
    Fn.ops.gemm_rowln = spy_rowln
    f1 = br(batch["vis_fea"], batch["vis_fea_mask"], batch["q_ipt"], batch["q_ipt_graph"], batch["q_ipt_mask"], True)
print("decoder output fused vs chain:", rel(f1, f0), "rowln calls", len(spy))
# spy order per layer: [Wv+LN (mode1,K=512), Wq (mode0), W1 (mode0,N=2048), W2+LN (mode1,K=2048)]
for i in range(6):
    s_self, s_q, s_w1, s_w2 = spy[4 * i:4 * i + 4]
    print(f"layer {i}: self-attn y {rel(s_self[3]['y'], taps['dec_self_attention_%d' % i].reshape(128, -1)):.3e}  "
          f"ffn y {rel(s_w2[3]['y'], taps['dec_feed_forward_%d' % i].reshape(128, -1)):.3e}  modes {s_self[0]}{s_q[0]}{s_w1[0]}{s_w2[0]}")
