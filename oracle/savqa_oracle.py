"""CPU oracle for the SA-VQA graph-guided attention encoder path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch CPU restatement (PyTorch CPU tensors, fp32 or fp64) of the
arithmetic that the reference performs on the hot path named by BASELINE.json:

    reference/models/modules.py      (embedding, layer_normalization, multihead_attention,
                                      new_multihead_attention, new_multihead_attention_with_graph_mask,
                                      feedforward, label_smoothing)
    reference/models/AttModel_x3.py  (AttModel_vis_grid.forward, AttModel_syb.forward, classifier heads)
    reference/models/main_itp_ddp_tar_super_node.py:335-345 (label-smoothed 3-head loss)

Why torch-on-CPU and not numpy/C: the arithmetic of the reference lives in a third-party
dependency -- PyTorch (unpinned in the reference: no requirements file; container has 2.11.0) --
and the reference's backward pass *is* torch.autograd.  Restating the forward with CPU torch ops
gives the gradient oracle for free and keeps the matmul/softmax rounding identical to what the
reference executes on a CPU.  There is no C restatement to compile (build() has nothing to do here).

Pinning: the reference has NO tests, golden vectors or fixtures of its own (SURVEY.md section 4,
8(c)), so the oracle is pinned against the *live* reference modules, imported unmodified from
/root/reference in the authoring container by oracle/make_golden.py, which commits seeded
input/output vectors under tests/golden/.  tests/test_oracle_golden.py checks this file against
those vectors (<= 2e-6 norm-rel in fp32, masks/gathers bit-exact); re-running oracle/make_golden.py
where /root/reference is present regenerates the fixtures bit for bit.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs (cpu_baseline, --impl reference, and the
stock-PyTorch-on-the-same-GPU comparator --impl stock-gpu, where these functions run on CUDA tensors through
ATen / cuBLAS) may import this module.  The product package (structured-alignment-vqa_b200/savqa_b200) never does.

Every function takes plain tensors and parameters keyed by the reference's state_dict names.
`operand_dtype=torch.bfloat16` emulates "bf16 MMA operands, fp32 accumulate" (used only to calibrate
the per-tensor tolerances stated in the GPU parity tests).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

#: masking constant of the reference, `-2 ** 32 + 1` (modules.py:167,261,369); rounds to -4294967296.0f in fp32
KEY_MASK_FILL = float(-2 ** 32 + 1)
PAD_WORD = 400000  # AttModel_x3.py:13


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def _mm_operand(t: Tensor, operand_dtype: Optional[torch.dtype]) -> Tensor:
    """Round a matmul operand to `operand_dtype` and come back (bf16-operand emulation)."""
    if operand_dtype is None:
        return t
    return t.to(operand_dtype).to(t.dtype)


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], operand_dtype=None) -> Tensor:
    """nn.Linear: y = x W^T + b (modules.py:227-229, 428-429)."""
    y = _mm_operand(x, operand_dtype) @ _mm_operand(weight, operand_dtype).transpose(-1, -2)
    return y if bias is None else y + bias


def sub(params: Params, prefix: str) -> Params:
    """Slice of a state_dict below `prefix.` (prefix stripped)."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in params.items() if k.startswith(p)}


# --------------------------------------------------------------------------------------------
# modules.py primitives
# --------------------------------------------------------------------------------------------
def embedding_lookup(idx: Tensor, table: Tensor, scale: bool) -> Tensor:
    """modules.py:32-46 `embedding.forward`: row gather, optionally times sqrt(num_units).

    (`zeros_pad` only changes which row gets no gradient: padding_idx 0 or -1 -> last row.)"""
    out = table[idx]
    if scale:
        out = out * (table.shape[1] ** 0.5)
    return out


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-8) -> Tensor:
    """modules.py:62-65: gamma*(x-mean)/(std_unbiased+eps)+beta  (eps added to sigma, n-1 divisor)."""
    mu = x.mean(-1, keepdim=True)
    # Tensor.std (unbiased) on purpose: its autograd formula returns 0 -- not NaN -- for a sigma == 0 row, and
    # that is the reference's backward semantics for constant rows (dx = (g - mean g) / eps there).
    sigma = x.std(-1, keepdim=True)
    return gamma * (x - mu) / (sigma + eps) + beta


def _heads(x: Tensor, num_heads: int) -> Tensor:
    """cat(chunk(x, H, dim=2), dim=0): [N,T,C] -> [H*N,T,C/H], row index h*N+n (modules.py:246-248)."""
    n, t, c = x.shape
    return x.reshape(n, t, num_heads, c // num_heads).permute(2, 0, 1, 3).reshape(num_heads * n, t, c // num_heads)


def _merge_heads(x: Tensor, num_heads: int) -> Tensor:
    """Inverse of _heads (modules.py:301)."""
    hn, t, d = x.shape
    n = hn // num_heads
    return x.reshape(num_heads, n, t, d).permute(1, 2, 0, 3).reshape(n, t, num_heads * d)


def attention(
    queries: Tensor,
    keys: Tensor,
    values: Tensor,
    graph: Optional[Tensor],
    p: Params,
    num_heads: int,
    causality: bool = False,
    renorm: str = "l1clamp",
    operand_dtype=None,
) -> Tuple[Tensor, Tensor]:
    """The three attention modules of modules.py in one restatement.

    renorm="l1clamp": new_multihead_attention (modules.py:236-311): W = A / max(sum|A|, 1e-12), A = G*softmax(S)
    renorm="addeps" : new_multihead_attention_with_graph_mask (modules.py:339-403): W = A / (sum A + 1e-7)
    renorm="none"   : multihead_attention (modules.py:143-207): graph is None, W = softmax(S)

    Returns (layer-normed output [N,Tq,C], attention weights [H*N,Tq,Tk] BEFORE the query mask).
    """
    Q = torch.relu(linear(queries, p["Q_proj.0.weight"], p["Q_proj.0.bias"], operand_dtype))
    K = torch.relu(linear(keys, p["K_proj.0.weight"], p["K_proj.0.bias"], operand_dtype))
    V = torch.relu(linear(values, p["V_proj.0.weight"], p["V_proj.0.bias"], operand_dtype))
    Qh, Kh, Vh = _heads(Q, num_heads), _heads(K, num_heads), _heads(V, num_heads)
    d = Kh.shape[-1]

    # scores, divided AFTER the contraction (modules.py:251-254)
    S = torch.bmm(_mm_operand(Qh, operand_dtype), _mm_operand(Kh, operand_dtype).transpose(1, 2)) / (d ** 0.5)

    # key padding mask from the RAW key inputs: sign(|sum_c keys|) (modules.py:257-263)
    key_on = (keys.sum(-1) != 0)  # [N,Tk]
    key_on = key_on.repeat(num_heads, 1).unsqueeze(1)  # [H*N,1,Tk]
    fill = torch.full((), KEY_MASK_FILL, dtype=S.dtype, device=S.device)
    S = torch.where(key_on, S, fill)
    if causality:  # modules.py:268-275
        tq, tk = S.shape[1], S.shape[2]
        keep = torch.ones(tq, tk, dtype=torch.bool, device=S.device).tril()
        S = torch.where(keep, S, fill)

    P = torch.softmax(S, dim=-1)  # modules.py:278 (max-subtracted over ALL keys)
    if renorm == "none":
        W = P
    else:
        A = graph.to(P.dtype).repeat(num_heads, 1, 1) * P  # modules.py:281-284
        if renorm == "l1clamp":
            W = A / A.abs().sum(-1, keepdim=True).clamp_min(1e-12)  # F.normalize(p=1), modules.py:285
        elif renorm == "addeps":
            W = A / (A.sum(-1, keepdim=True) + 1e-7)  # modules.py:378
        else:
            raise ValueError(renorm)
    att = W

    # query mask from the RAW query inputs (modules.py:289-292)
    q_on = (queries.sum(-1) != 0).to(W.dtype).repeat(num_heads, 1).unsqueeze(2)  # [H*N,Tq,1]
    Wq = W * q_on

    O = torch.bmm(_mm_operand(Wq, operand_dtype), _mm_operand(Vh, operand_dtype))  # modules.py:298
    out = _merge_heads(O, num_heads) + queries  # residual adds the RAW queries (modules.py:304)
    out = layer_norm(out, p["normalization.gamma"], p["normalization.beta"])
    return out, att


def feedforward(x: Tensor, p: Params, operand_dtype=None) -> Tensor:
    """modules.py:432-447 (Linear branch): LN(W2 relu(W1 x + b1) + b2 + x)."""
    h = torch.relu(linear(x, p["conv1.0.weight"], p["conv1.0.bias"], operand_dtype))
    z = linear(h, p["conv2.weight"], p["conv2.bias"], operand_dtype) + x
    return layer_norm(z, p["normalization.gamma"], p["normalization.beta"])


def label_smoothing(x: Tensor, epsilon: float = 0.1) -> Tensor:
    """modules.py:460-463."""
    return (1 - epsilon) * x + epsilon / x.shape[-1]


# --------------------------------------------------------------------------------------------
# AttModel_x3.py: scene-graph mask construction
# --------------------------------------------------------------------------------------------
def build_masks(
    first_mask: Tensor, q_mask: Tensor, q_graph: Tensor, first_graph: Optional[Tensor], dec_mask_on: bool,
    dtype=torch.float32,
) -> Tuple[Tensor, Tensor, Tensor]:
    """AttModel_x3.py:103-122 (vis: first_graph=None -> top-left block all ones) and :229-247 (syb).

    first_mask [B,V,V], q_mask [B,Q,Q], q_graph [B,Q,Q], first_graph [B,V,V] | None
    returns graph_diag [B,T,T], graph [B,T,T], dec_mask [B,1,T] (T = V+Q), all `dtype`.

    The reference builds `mask = block_diag(first_mask, q_mask)` per sample in a Python loop and then
    aliases graph_cross and graph (`graph = graph_cross`, in-place writes), so blocks 2..5 of the
    encoder all see the same final `graph`:  TL = ones | first_graph, TR = BL = 1 - 0 = 1, BR = q_graph.
    """
    B, V = first_mask.shape[0], first_mask.shape[1]
    Qn = q_mask.shape[1]
    T = V + Qn
    graph_diag = torch.zeros(B, T, T, dtype=dtype, device=first_mask.device)
    graph_diag[:, V:, V:] = q_mask.to(dtype)
    graph = torch.ones(B, T, T, dtype=dtype, device=first_mask.device)
    if first_graph is not None:
        graph[:, :V, :V] = first_graph.to(dtype)
    graph[:, V:, V:] = q_graph.to(dtype)
    dec_mask = torch.zeros(B, 1, T, dtype=dtype, device=first_mask.device)
    if dec_mask_on:
        # row sums of the block-diagonal mask; rows that are != 0 become 1 (AttModel_x3.py:113-116)
        dec_mask[:, 0, :V] = (first_mask.to(dtype).sum(-1) != 0).to(dtype)
        dec_mask[:, 0, V:] = (q_mask.to(dtype).sum(-1) != 0).to(dtype)
    return graph_diag, graph, dec_mask


# --------------------------------------------------------------------------------------------
# AttModel_x3.py: the two branch models
# --------------------------------------------------------------------------------------------
def branch_forward(
    params: Params,
    kind: str,
    first_ipt: Tensor,
    first_mask: Tensor,
    first_graph: Optional[Tensor],
    q_ipt: Tensor,
    q_graph: Tensor,
    q_mask: Tensor,
    dec_mask_on: bool,
    num_blocks: int,
    num_heads: int,
    operand_dtype=None,
    taps: Optional[dict] = None,
) -> Tensor:
    """AttModel_vis_grid.forward (AttModel_x3.py:91-156, kind="vis") and AttModel_syb.forward
    (:214-282, kind="syb") with dropout off.  `params` uses the branch's own state_dict key names.
    Returns dec [B,1,C].  `taps`, if given, receives intermediate tensors by name.
    """
    assert kind in ("vis", "syb")
    dt = params["syb_mlp2.weight"].dtype
    if first_ipt.dim() == 4:  # grid features [B,gx,gy,F] (AttModel_x3.py:94-95)
        first_ipt = first_ipt.reshape(first_ipt.shape[0], -1, first_ipt.shape[3])
    first_ipt = first_ipt.to(dt)
    B, V = first_ipt.shape[0], first_ipt.shape[1]

    q = params["syb_emb.weight"][q_ipt]  # :96 / :216
    q = torch.relu(linear(q, params["syb_mlp.0.weight"], params["syb_mlp.0.bias"], operand_dtype))  # :97
    x = torch.cat([first_ipt, q], dim=1)  # :98
    x = linear(x, params["syb_mlp2.weight"], params["syb_mlp2.bias"], operand_dtype)  # :99
    T = x.shape[1]
    pos_key = "syb_positional_encoding.0.lookup_table" if kind == "vis" else "syb_positional_encoding.lookup_table"
    x = x + params[pos_key][:T].unsqueeze(0)  # :100-101 (scale=False)
    if taps is not None:
        taps["enc_input"] = x

    graph_diag, graph, dec_mask = build_masks(first_mask, q_mask, q_graph, first_graph, dec_mask_on, dtype=dt)
    if taps is not None:
        taps["graph_diag"], taps["graph"], taps["dec_mask"] = graph_diag, graph, dec_mask

    for i in range(num_blocks):  # :127-139; blocks 0,1 see graph_diag, the rest the aliased graph
        g = graph_diag if i < 2 else graph
        x, att = attention(x, x, x, g, sub(params, f"enc_self_attention_{i}"), num_heads, operand_dtype=operand_dtype)
        if taps is not None:
            taps[f"enc_att_{i}"], taps[f"enc_attw_{i}"] = x, att
        x = feedforward(x, sub(params, f"enc_feed_forward_{i}"), operand_dtype)
        if taps is not None:
            taps[f"enc_ffn_{i}"] = x
    memory = x

    # decoder input: class-token row 2 of dec_emb times sqrt(C), plus position 0 (:141-147)
    C = memory.shape[-1]
    dec = params["dec_emb.lookup_table"][2] * (C ** 0.5) + params["dec_positional_encoding.lookup_table"][0]
    dec = dec.reshape(1, 1, C).repeat(B, 1, 1)
    for i in range(num_blocks):  # :148-154
        dec, _ = attention(dec, dec, dec, None, sub(params, f"dec_self_attention_{i}"), num_heads,
                           causality=True, renorm="none", operand_dtype=operand_dtype)
        dec, att = attention(dec, memory, memory, dec_mask, sub(params, f"dec_vanilla_attention_{i}"), num_heads,
                             operand_dtype=operand_dtype)
        dec = feedforward(dec, sub(params, f"dec_feed_forward_{i}"), operand_dtype)
        if taps is not None:
            taps[f"dec_{i}"], taps[f"dec_attw_{i}"] = dec, att
    return dec


# --------------------------------------------------------------------------------------------
# classifier heads + loss
# --------------------------------------------------------------------------------------------
def mlp_head(x: Tensor, p: Params, operand_dtype=None) -> Tensor:
    """Linear + ReLU + Dropout(off) + Linear (AttModel_x3.py:482-500)."""
    h = torch.relu(linear(x, p["0.weight"], p["0.bias"], operand_dtype))
    return linear(h, p["3.weight"], p["3.bias"], operand_dtype)


def answer_logits(params: Params, fea_vis: Tensor, fea_syb: Tensor, operand_dtype=None):
    """AttModel_x3.py:531-541 with mcb=False: (logits_concat, logits_vis, logits_syb)."""
    logits_vis = mlp_head(fea_vis, sub(params, "cls_vis"), operand_dtype).squeeze(1)
    logits_syb = mlp_head(fea_syb, sub(params, "cls_syb"), operand_dtype).squeeze(1)
    fea = torch.cat((fea_syb.squeeze(1), fea_vis.squeeze(1)), 1)  # (reference .squeeze() also drops B when B == 1)
    logits_concat = mlp_head(fea, sub(params, "cls"), operand_dtype)
    return logits_concat, logits_vis, logits_syb


def answer_loss(logits_concat: Tensor, logits_vis: Tensor, logits_syb: Tensor, answer: Tensor, epsilon: float = 0.1):
    """main_itp_ddp_tar_super_node.py:335-345: mean of three log-softmaxes against a label-smoothed one-hot."""
    lsm = (torch.log_softmax(logits_vis, -1) + torch.log_softmax(logits_syb, -1) + torch.log_softmax(logits_concat, -1)) / 3
    one_hot = torch.zeros_like(logits_concat)
    one_hot.scatter_(1, answer.view(-1, 1), 1)
    target = label_smoothing(one_hot, epsilon)
    return -(target * lsm).sum(-1).mean()


def encoder_step(
    params: Params,
    batch: Dict[str, Tensor],
    num_blocks: int,
    num_heads: int,
    dec_mask_on: bool = True,
    operand_dtype=None,
):
    """One pass of the hot path as the train loop drives it (main...:321-345) WITHOUT MIL_NCE:
    `batch["syb_ipt"]` [B,M,2048] stands for MIL_NCE's `new_macro_ipt` output (AttModel_x3.py:525, 530).
    Returns (loss, (logits_concat, logits_vis, logits_syb), fea_vis, fea_syb)."""
    fea_vis = branch_forward(sub(params, "att_vis_grid"), "vis", batch["vis_fea"], batch["vis_fea_mask"], None,
                             batch["q_ipt"], batch["q_ipt_graph"], batch["q_ipt_mask"], dec_mask_on,
                             num_blocks, num_heads, operand_dtype)
    fea_syb = branch_forward(sub(params, "att_syb"), "syb", batch["syb_ipt"], batch["macro_node_mask"],
                             batch["macro_graph_ipt"], batch["q_ipt"], batch["q_ipt_graph"], batch["q_ipt_mask"],
                             dec_mask_on, num_blocks, num_heads, operand_dtype)
    logits = answer_logits(params, fea_vis, fea_syb, operand_dtype)
    loss = answer_loss(*logits, batch["answer"])
    return loss, logits, fea_vis, fea_syb


# --------------------------------------------------------------------------------------------
# AttModel_x3.py:285-443  MIL_NCE (only_obj=True) and the full 16-argument step
# --------------------------------------------------------------------------------------------
def mil_nce(params: Params, vis_fea: Tensor, macro_ipt: Tensor, macro_obj_loc: Tensor, pos_ids: Tensor, neg_ids: Tensor,
            obj_mask: Tensor, operand_dtype=None):
    """MIL_NCE.forward with only_obj=True (AttModel_x3.py:339-379, 441-443).  `params` uses MIL_NCE's own key names.
    Returns (macro_ipt_output [B,M,2048], mil_nce_obj scalar)."""
    eps = 1e-6
    table = params["syb_emb.weight"]
    relu = torch.relu
    nodes = relu(linear(table[macro_ipt], params["marco_mlp.0.weight"], params["marco_mlp.0.bias"], operand_dtype)).detach().clone()  # :352-354
    pos = relu(linear(table[pos_ids], params["syb_mlp.0.weight"], params["syb_mlp.0.bias"], operand_dtype))    # [B,V,topN,h]  :356-357
    neg = relu(linear(table[neg_ids], params["syb_mlp.0.weight"], params["syb_mlp.0.bias"], operand_dtype))    # :358-359
    vis = relu(linear(vis_fea, params["vis_mlp.0.weight"], params["vis_mlp.0.bias"], operand_dtype))           # [B,V,h]  :361
    m4 = obj_mask.unsqueeze(3).to(vis.dtype)
    raw_pos = torch.matmul(_mm_operand(pos, operand_dtype), _mm_operand(vis, operand_dtype).unsqueeze(3))      # [B,V,topN,1]  :365
    raw_neg = torch.matmul(_mm_operand(neg, operand_dtype), _mm_operand(vis, operand_dtype).unsqueeze(3))      # :366
    s_pos = (m4 * raw_pos).clamp(min=eps)
    s_neg = (m4 * raw_neg).clamp(min=eps)
    floor = torch.zeros_like(s_neg).clamp(min=eps)
    obj = torch.mean(torch.logsumexp(torch.cat((s_pos, floor), dim=1), dim=2)
                     - torch.logsumexp(torch.cat((s_pos, s_neg), dim=1), dim=2))                               # :367
    refined = torch.sum(torch.softmax(raw_pos, dim=2) * pos, dim=2)                                             # [B,V,h]  :372-374
    b_idx, v_idx = (macro_obj_loc >= 0).nonzero(as_tuple=True)
    nodes[b_idx, macro_obj_loc[b_idx, v_idx].long(), :] = refined[b_idx, v_idx, :]                              # :377-379
    out = relu(linear(nodes, params["ipt_mlp.0.weight"], params["ipt_mlp.0.bias"], operand_dtype))              # :441
    return out, obj


def full_step(params: Params, batch: Dict[str, Tensor], num_blocks: int, num_heads: int, dec_mask_on: bool = True,
              with_milnce_loss: bool = True, operand_dtype=None):
    """One pass of the path as the train loop drives it WITH MIL_NCE (main...:321-345, 359-360; only_obj): the reference's
    16-argument AttModel.forward (AttModel_x3.py:512-542) + the loss (+ mil_nce_loss = -mil_nce_obj when with_MILNCE_loss).
    Returns (loss, (logits_concat, logits_vis, logits_syb), mil_nce_obj, new_macro_ipt)."""
    syb_ipt, obj = mil_nce(sub(params, "MIL_NCE"), batch["vis_fea"], batch["macro_node_ipt"], batch["macro_obj_loc_ipt"],
                           batch["micro_positive_obj_ipt"], batch["micro_negative_obj_ipt"], batch["micro_obj_mask"], operand_dtype)
    b2 = dict(batch, syb_ipt=syb_ipt)
    loss, logits, _, _ = encoder_step(params, b2, num_blocks, num_heads, dec_mask_on, operand_dtype)
    if with_milnce_loss:
        loss = loss - obj
    return loss, logits, obj, syb_ipt


def rel_err(a: Tensor, b: Tensor) -> float:
    """||a-b|| / ||b||  (the per-tensor parity metric of SURVEY.md 8(c))."""
    a = a.detach().double().reshape(-1)
    b = b.detach().double().reshape(-1)
    denom = float(torch.linalg.norm(b))
    num = float(torch.linalg.norm(a - b))
    return num / denom if denom > 0 else num
