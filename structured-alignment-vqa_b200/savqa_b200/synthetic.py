"""Seeded synthetic batches shaped like the reference's `collate_fn` output
(models/data_loader_itp_bbox_super_node_onlyobj.py:341-445) and model construction with the production
hyper-parameters.  Used by bench.py, __graft_entry__.smoke() and the GPU tests; no dataset or
checkpoint is needed (there is no network on the GPU box)."""
from __future__ import annotations

import types
from typing import Dict, Optional

import torch

from . import AttModel_x3 as A
from .modules import layer_normalization

#: BASELINE.json configs (V regions, Q question tokens, M symbolic nodes)
GQA_SHAPED = dict(name="gqa_shaped", hidden=512, heads=8, blocks=6, maxlen=300, maxlen_q=50, maxlen_v=49, ncls=1845,
                  hidden_mil=64, nrel=311, V=36, Q=20, M=108, topN=1)
CFG1 = dict(GQA_SHAPED, name="cfg1_cpu", M=60)
CFG2 = dict(GQA_SHAPED, name="cfg2_inference", V=100, Q=20, M=279, maxlen=300)
TINY = dict(name="tiny", hidden=64, heads=4, blocks=6, maxlen=40, maxlen_q=12, maxlen_v=9, ncls=20, hidden_mil=16, nrel=5,
            V=5, Q=4, M=7, topN=1)


def build_model(cfg: Dict, vocab_rows: Optional[int] = None, dropout: float = 0.0, randomize_ln: bool = True, seed: int = 0) -> A.AttModel:
    """Reference initialisation (nn.Linear Kaiming-uniform, Xavier-normal tables) with `torch.manual_seed(seed)`, then
    LN gamma~U(0.8,1.2), beta~N(0,0.1): with the default gamma=1, beta=0 the reference's activation-derived padding masks
    are decided by rounding noise.  `vocab_rows` shrinks the 407000-row word tables (tests)."""
    torch.manual_seed(seed)
    glove = types.SimpleNamespace(vectors=torch.randn(1000, 300))
    saved = A.VOCAB_ROWS
    if vocab_rows is not None:
        A.VOCAB_ROWS = vocab_rows
    try:
        model = A.AttModel(glove, cfg["hidden"], cfg["hidden_mil"], cfg["ncls"], cfg["maxlen_q"], cfg["maxlen"], cfg["maxlen_v"],
                           cfg["blocks"], cfg["heads"], dropout, 0.1, cfg["nrel"], True)
    finally:
        A.VOCAB_ROWS = saved
    if randomize_ln:
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, layer_normalization):
                    m.beta.normal_(0, 0.1)
                    m.gamma.uniform_(0.8, 1.2)
    return model


def make_batch(cfg: Dict, batch_size: int, seed: int = 0, vocab_rows: Optional[int] = None, device="cpu", full_length: bool = False,
               pin: bool = False) -> Dict[str, torch.Tensor]:
    """One collate_fn-shaped batch.  Ragged valid lengths (sample 0 full), zero padding rows, int32 masks and graphs,
    int64 word ids with PAD on padding.  `syb_ipt` [B,M,2048] stands for MIL_NCE's output (AttModel_x3.py:525-530)."""
    g = torch.Generator().manual_seed(seed)
    B, V, Q, M = batch_size, cfg["V"], cfg["Q"], cfg["M"]
    n_words = min(400000, (vocab_rows or A.VOCAB_ROWS) - 1)
    pad = min(A.PAD, (vocab_rows or A.VOCAB_ROWS) - 1)

    def lens(n):
        if full_length:
            return torch.full((B,), n, dtype=torch.int64)
        l = torch.randint(max(1, n // 2), n + 1, (B,), generator=g)
        l[0] = n
        return l

    vl, ql, ml = lens(V), lens(Q), lens(M)
    ar = torch.arange
    vvalid = ar(V)[None, :] < vl[:, None]
    qvalid = ar(Q)[None, :] < ql[:, None]
    mvalid = ar(M)[None, :] < ml[:, None]

    def block(valid):
        return (valid[:, :, None] & valid[:, None, :])

    vis_fea = torch.rand(B, V, 2048, generator=g) * vvalid[:, :, None]
    syb_ipt = torch.rand(B, M, 2048, generator=g) * mvalid[:, :, None]
    q_ipt = torch.randint(0, n_words, (B, Q), generator=g)
    q_ipt[~qvalid] = pad
    out = dict(
        vis_fea=vis_fea,
        vis_fea_mask=block(vvalid).int(),
        q_ipt=q_ipt,
        q_ipt_mask=block(qvalid).int(),
        q_ipt_graph=((torch.rand(B, Q, Q, generator=g) < 0.2) & block(qvalid)).int(),
        syb_ipt=syb_ipt,
        macro_node_ipt=torch.randint(0, n_words, (B, M), generator=g),
        macro_node_mask=block(mvalid).int(),
        macro_graph_ipt=((torch.rand(B, M, M, generator=g) < 0.1) & block(mvalid)).int(),
        answer=torch.randint(0, cfg["ncls"], (B,), generator=g),
    )
    # MIL_NCE inputs (only needed by the full AttModel.forward)
    topn = cfg["topN"]
    loc = torch.full((B, V), -1, dtype=torch.int64)
    for b in range(B):
        n = int(min(vl[b], ml[b]))
        loc[b, :n] = torch.randperm(int(ml[b]), generator=g)[:n]
    out.update(macro_obj_loc_ipt=loc,
               micro_positive_obj_ipt=torch.randint(0, n_words, (B, V, topn), generator=g),
               micro_negative_obj_ipt=torch.randint(0, n_words, (B, V, topn), generator=g),
               micro_obj_mask=vvalid[:, :, None].expand(B, V, topn).int().contiguous())
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    if str(device) != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


ENCODER_KEYS = ("vis_fea", "vis_fea_mask", "q_ipt", "q_ipt_mask", "q_ipt_graph", "syb_ipt", "macro_node_mask", "macro_graph_ipt", "answer")


def branch_flops(cfg: Dict, T: int, with_heads: bool = False) -> float:
    """Algorithmic forward FLOPs of one branch for one sample (matches torch's FlopCounterMode on the
    reference exactly): dense-equivalent, padding of MMA tiles never counted."""
    C, L, E, F_, Q = cfg["hidden"], cfg["blocks"], 300, 2048, cfg["Q"]
    f = 2 * E * F_ * Q + 2 * F_ * C * T
    f += L * T * 6 * C * C + L * 4 * T * T * C + L * T * 16 * C * C
    f += L * (6 * C * C + 4 * C) + L * 2 * C * C + L * T * 4 * C * C + L * 4 * T * C + L * 16 * C * C
    return float(f)


def mil_nce_flops(cfg: Dict) -> float:
    """Algorithmic forward FLOPs of MIL_NCE (only_obj) for one sample: marco_mlp, syb_mlp on the positive and negative words,
    vis_mlp, the object-word scores and the refinement, ipt_mlp (AttModel_x3.py:352-379, 441)."""
    E, F_, h, V, M, n = 300, 2048, cfg["hidden_mil"], cfg["V"], cfg["M"], cfg["topN"]
    return float(2 * E * h * M + 2 * E * h * 2 * V * n + 2 * F_ * h * V + 3 * 2 * V * n * h + 2 * h * F_ * M)


def step_flops(cfg: Dict, batch_size: int, backward: bool = True, full: bool = False) -> float:
    """fwd (+ bwd = 2x fwd) FLOPs of the step: both branches + the three classifier heads (+ MIL_NCE when `full`)."""
    C, ncls = cfg["hidden"], cfg["ncls"]
    per = branch_flops(cfg, cfg["V"] + cfg["Q"]) + branch_flops(cfg, cfg["M"] + cfg["Q"])
    per += 2 * (2 * C * C + 2 * C * ncls) + (4 * C * C + 2 * C * ncls)
    if full:
        per += mil_nce_flops(cfg)
    return per * batch_size * (3.0 if backward else 1.0)
