"""Deterministic inputs / weights shared by oracle/make_golden.py (which runs the LIVE reference in the
authoring container) and the tests (which re-create the same tensors anywhere and compare against the
committed outputs in tests/golden/).  TEST INFRASTRUCTURE ONLY.

Nothing big is stored in the fixtures: every weight and input is regenerated from a name-keyed seed
(torch CPU mt19937 `randn`/`rand` are bit-stable for a fixed torch version; each fixture also stores an
fp64 checksum of what the generator produced so that RNG drift fails loudly instead of silently).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
SMALL_VOCAB = 64  # golden cases only ever index rows [0, 64) of the 407000x300 word tables
E_GLOVE = 300
F_REGION = 2048


def _gen(name: str) -> torch.Generator:
    seed = int.from_bytes(hashlib.sha256(name.encode()).digest()[:7], "little")
    return torch.Generator().manual_seed(seed)


def randn(name: str, *shape, scale: float = 1.0) -> Tensor:
    return torch.randn(*shape, generator=_gen(name)) * scale


def rand(name: str, *shape, lo: float = 0.0, hi: float = 1.0) -> Tensor:
    return torch.rand(*shape, generator=_gen(name)) * (hi - lo) + lo


def randint(name: str, lo: int, hi: int, *shape) -> Tensor:
    return torch.randint(lo, hi, shape, generator=_gen(name))


def bernoulli(name: str, p: float, *shape) -> Tensor:
    return (torch.rand(*shape, generator=_gen(name)) < p)


def checksum(tensors: Dict[str, Tensor]) -> float:
    tot = 0.0
    for k in sorted(tensors):
        t = tensors[k].double()
        tot += float(t.abs().sum()) + 3.0 * float(t.sum())
    return tot


BIG = 20000  # tensors above this many elements are stored as two seeded projections + norm


def fingerprint(name: str, g: Tensor) -> Dict[str, Tensor]:
    """Compact, order-sensitive summary of a 2-D gradient that is too big to commit in full:
    g @ r_in [out], r_out @ g [in] with name-seeded Gaussian r, plus (norm, sum)."""
    g = g.detach().double()
    if g.dim() != 2 or g.numel() <= BIG:
        return {"full": g.float()}
    r_in = randn(f"fp/{name}/in", g.shape[1]).double()
    r_out = randn(f"fp/{name}/out", g.shape[0]).double()
    return {"rows": (g @ r_in).float(), "cols": (r_out @ g).float(),
            "norm": torch.tensor([float(g.norm()), float(g.sum())])}


def pack_grads(case: str, grads: Dict[str, Tensor]) -> Dict[str, Tensor]:
    out = {}
    for k, g in grads.items():
        for part, v in fingerprint(f"{case}/{k}", g).items():
            out[f"grad/{k}/{part}"] = v
    return out


def compare_grads(case: str, got: Dict[str, Tensor], golden, rel_tol: float, keys=None) -> Dict[str, float]:
    """Return {param: worst norm-rel error over the stored parts}; raises AssertionError above rel_tol."""
    errs = {}
    for k, g in got.items():
        if keys is not None and k not in keys:
            continue
        fp = fingerprint(f"{case}/{k}", g)
        worst = 0.0
        for part, v in fp.items():
            ref = torch.from_numpy(golden[f"grad/{k}/{part}"]).double()
            v = v.double()
            den = float(ref.norm())
            err = float((v - ref).norm()) / den if den > 0 else float((v - ref).norm())
            worst = max(worst, err)
        errs[k] = worst
        assert worst <= rel_tol, f"{case}: grad {k} rel err {worst:.3e} > {rel_tol:.1e}"
    return errs


# --------------------------------------------------------------------------------------------
# parameter-name contract; make_golden.py asserts the live reference agrees
# --------------------------------------------------------------------------------------------
def attention_shapes(C: int) -> List[Tuple[str, Tuple[int, ...]]]:
    out = []
    for nm in ("Q_proj", "K_proj", "V_proj"):
        out += [(f"{nm}.0.weight", (C, C)), (f"{nm}.0.bias", (C,))]
    out += [("normalization.gamma", (C,)), ("normalization.beta", (C,))]
    return out


def feedforward_shapes(C: int, hidden: Optional[int] = None) -> List[Tuple[str, Tuple[int, ...]]]:
    hidden = hidden or 4 * C
    return [("conv1.0.weight", (hidden, C)), ("conv1.0.bias", (hidden,)),
            ("conv2.weight", (C, hidden)), ("conv2.bias", (C,)),
            ("normalization.gamma", (C,)), ("normalization.beta", (C,))]


def branch_shapes(kind: str, C: int, maxlen: int, maxlen_q: int, maxlen_v: int, num_blocks: int, ncls: int,
                  vocab: int = 407000) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict keys/shapes of AttModel_vis_grid (AttModel_x3.py:27-90) / AttModel_syb (:160-212), in
    registration order."""
    L: List[Tuple[str, Tuple[int, ...]]] = []

    def enc_blocks():
        for i in range(num_blocks):
            for k, s in attention_shapes(C):
                L.append((f"enc_self_attention_{i}.{k}", s))
            for k, s in feedforward_shapes(C):
                L.append((f"enc_feed_forward_{i}.{k}", s))

    def dec_blocks():
        for i in range(num_blocks):
            for k, s in attention_shapes(C):
                L.append((f"dec_self_attention_{i}.{k}", s))
            for k, s in attention_shapes(C):
                L.append((f"dec_vanilla_attention_{i}.{k}", s))
            for k, s in feedforward_shapes(C):
                L.append((f"dec_feed_forward_{i}.{k}", s))

    L += [("syb_emb.weight", (vocab, E_GLOVE)),
          ("syb_mlp.0.weight", (F_REGION, E_GLOVE)), ("syb_mlp.0.bias", (F_REGION,)),
          ("syb_mlp2.weight", (C, F_REGION)), ("syb_mlp2.bias", (C,))]
    if kind == "vis":
        L += [("v_mlp.0.weight", (C, F_REGION)), ("v_mlp.0.bias", (C,)), ("v_mlp.2.weight", (C, C)), ("v_mlp.2.bias", (C,)),
              ("v_positional_encoding.0.lookup_table", (maxlen_v, C)),
              ("input_proj.weight", (C, F_REGION)), ("input_proj.bias", (C,))]
        enc_blocks()
        L += [("q_mlp.0.weight", (C, E_GLOVE)), ("q_mlp.0.bias", (C,)), ("q_mlp.2.weight", (C, C)), ("q_mlp.2.bias", (C,)),
              ("q_positional_encoding.0.lookup_table", (maxlen_q, C)),
              ("syb_positional_encoding.0.lookup_table", (maxlen, C)),
              ("dec_emb.lookup_table", (ncls, C)),
              ("dec_positional_encoding.lookup_table", (maxlen, C))]
        dec_blocks()
    else:
        L += [("syb_positional_encoding.lookup_table", (maxlen + maxlen_q, C)),
              ("q_mlp.0.weight", (C, E_GLOVE)), ("q_mlp.0.bias", (C,)), ("q_mlp.1.weight", (C, C)), ("q_mlp.1.bias", (C,)),
              ("q_positional_encoding.0.lookup_table", (maxlen_q, C)),
              ("dec_emb.lookup_table", (ncls, C)),
              ("dec_positional_encoding.lookup_table", (maxlen + maxlen_q, C))]
        dec_blocks()
        enc_blocks()
    return L


def head_shapes(C: int, ncls: int) -> List[Tuple[str, Tuple[int, ...]]]:
    L = []
    for nm, cin in (("cls", 2 * C), ("cls_vis", C), ("cls_syb", C)):
        L += [(f"{nm}.0.weight", (C, cin)), (f"{nm}.0.bias", (C,)), (f"{nm}.3.weight", (ncls, C)), (f"{nm}.3.bias", (ncls,))]
    return L


def make_params(case: str, shapes: List[Tuple[str, Tuple[int, ...]]], small_vocab: int = SMALL_VOCAB) -> Dict[str, Tensor]:
    """Name-keyed deterministic parameters.  LN affine is randomised (gamma~U(0.8,1.2), beta~N(0,0.1)) because
    with the default gamma=1, beta=0 the reference's activation-derived key/query masks are decided by
    rounding noise.  Word tables are generated at `small_vocab` rows."""
    P: Dict[str, Tensor] = {}
    for k, s in shapes:
        nm = f"{case}/{k}"
        if k.endswith("gamma"):
            P[k] = rand(nm, *s, lo=0.8, hi=1.2)
        elif k.endswith("beta"):
            P[k] = randn(nm, *s, scale=0.1)
        elif k.endswith("bias"):
            P[k] = randn(nm, *s, scale=0.1)
        elif k == "syb_emb.weight":
            P[k] = randn(nm, small_vocab, s[1], scale=0.5)
        elif k.endswith("lookup_table"):
            P[k] = randn(nm, *s, scale=0.1)
        else:  # Linear weight [out, in]
            P[k] = randn(nm, *s, scale=1.0 / (s[1] ** 0.5))
    return P


# --------------------------------------------------------------------------------------------
# golden cases
# --------------------------------------------------------------------------------------------
def attention_case(case: str, C: int, N: int, Tq: int, Tk: int, self_att: bool, graph_p: float = 0.4):
    """Inputs for one attention-module call, with the edge cases the path has: an all-zero graph row, an
    all-zero key row (key mask), an all-zero query row (query mask), a row whose only graph-allowed key is
    key-masked (exercises the 1e-12 clamp)."""
    q = randn(f"{case}/queries", N, Tq, C)
    k = q if self_att else randn(f"{case}/keys", N, Tk, C)
    graph = bernoulli(f"{case}/graph", graph_p, N, Tq, Tk).float()
    if Tk >= 4:
        k = k.clone()
        k[0, Tk - 1, :] = 0  # padded key
        if self_att:
            q = k
        if Tq >= 3:
            graph[0, 1, :] = 0  # query row without any edge -> all-zero attention row
            graph[0, 2, :] = 0
            graph[0, 2, Tk - 1] = 1  # only edge goes to the masked key -> clamp path
        if not self_att and Tq >= 2:
            q = q.clone()
            q[N - 1, Tq - 1, :] = 0  # padded query
    return q, k, graph


def branch_case(case: str, kind: str, B: int, V: int, Q: int, small_vocab: int = SMALL_VOCAB):
    """A collate_fn-shaped mini batch (data_loader_itp_bbox_super_node_onlyobj.py:341-445) for one branch:
    ragged valid lengths, zero padding rows, int32 masks / graphs, int64 word ids."""
    v_len = randint(f"{case}/v_len", max(1, V // 2), V + 1, B)
    q_len = randint(f"{case}/q_len", max(1, Q // 2), Q + 1, B)
    v_len[0], q_len[0] = V, Q
    first = rand(f"{case}/first", B, V, F_REGION)
    first_mask = torch.zeros(B, V, V, dtype=torch.int32)
    q_mask = torch.zeros(B, Q, Q, dtype=torch.int32)
    q_graph = torch.zeros(B, Q, Q, dtype=torch.int32)
    first_graph = torch.zeros(B, V, V, dtype=torch.int32)
    q_ipt = torch.full((B, Q), small_vocab - 1, dtype=torch.int64)  # stands for PAD=400000
    qg = bernoulli(f"{case}/q_graph", 0.35, B, Q, Q)
    fg = bernoulli(f"{case}/first_graph", 0.3, B, V, V)
    ids = randint(f"{case}/q_ids", 0, small_vocab - 1, B, Q)
    for b in range(B):
        v, q = int(v_len[b]), int(q_len[b])
        first[b, v:] = 0
        first_mask[b, :v, :v] = 1
        q_mask[b, :q, :q] = 1
        q_graph[b, :q, :q] = qg[b, :q, :q].int()
        first_graph[b, :v, :v] = fg[b, :v, :v].int()
        q_ipt[b, :q] = ids[b, :q]
    return dict(first=first, first_mask=first_mask, first_graph=(first_graph if kind == "syb" else None),
                q_ipt=q_ipt, q_graph=q_graph, q_mask=q_mask)


# hyper-parameters of the small end-to-end golden model
SMALL = dict(C=64, heads=4, blocks=6, maxlen=40, maxlen_q=12, maxlen_v=9, ncls=20, B=3, V=5, Q=4, M=7)
# module-level goldens at the production width
WIDE = dict(C=512, heads=8)


# --------------------------------------------------------------------------------------------
# MIL_NCE (AttModel_x3.py:285-443, only_obj=True)
# --------------------------------------------------------------------------------------------
def mil_nce_shapes(h: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameters the only_obj forward uses (R, bilinear, rel_mlp stay at their initial values: never read)."""
    return [("syb_emb.weight", (407000, E_GLOVE)), ("marco_mlp.0.weight", (h, E_GLOVE)), ("marco_mlp.0.bias", (h,)),
            ("syb_mlp.0.weight", (h, E_GLOVE)), ("syb_mlp.0.bias", (h,)), ("vis_mlp.0.weight", (h, F_REGION)), ("vis_mlp.0.bias", (h,)),
            ("ipt_mlp.0.weight", (F_REGION, h)), ("ipt_mlp.0.bias", (F_REGION,))]


MIL_CASES = {"mil_nce_h16_top2": dict(h=16, topN=2, B=3, V=5, M=7), "mil_nce_h64_top1": dict(h=64, topN=1, B=4, V=6, M=9),
             "mil_nce_h128_top5": dict(h=128, topN=5, B=2, V=4, M=8)}


def mil_nce_case(case: str, B: int, V: int, M: int, topN: int, small_vocab: int = SMALL_VOCAB):
    """collate_fn-shaped MIL_NCE inputs (data_loader_itp_bbox_super_node_onlyobj.py:385-400): ragged object counts, distinct node
    positions per object (-1 padding), PAD-like ids and a zero mask on padded objects."""
    n_obj = randint(f"{case}/n_obj", max(1, V // 2), V + 1, B)
    n_obj[0] = V
    vis = rand(f"{case}/vis", B, V, F_REGION)
    macro_ipt = randint(f"{case}/macro_ipt", 0, small_vocab - 1, B, M)
    pos = randint(f"{case}/pos", 0, small_vocab - 1, B, V, topN)
    neg = randint(f"{case}/neg", 0, small_vocab - 1, B, V, topN)
    mask = torch.zeros(B, V, topN, dtype=torch.int32)
    loc = torch.full((B, V), -1, dtype=torch.int64)
    for b in range(B):
        n = int(n_obj[b])
        vis[b, n:] = 0
        pos[b, n:] = small_vocab - 1
        neg[b, n:] = small_vocab - 1
        mask[b, :n] = 1
        k = min(n, M)
        loc[b, :k] = torch.randperm(M, generator=_gen(f"{case}/loc{b}"))[:k]
    return dict(vis_fea=vis, macro_ipt=macro_ipt, macro_obj_loc=loc, pos=pos, neg=neg, mask=mask)
