"""Pins oracle/savqa_oracle.py against the golden vectors produced by the LIVE reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import golden_spec as GS
from oracle import savqa_oracle as O

FP32_TOL = 2e-6  # ||a-b||/||b||; same arithmetic, different op order only


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def t(a):
    return torch.from_numpy(np.asarray(a))


def check_checksum(g, tensors):
    got = GS.checksum(tensors)
    ref = float(g["checksum"])
    assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref)), "seeded generator drifted: regenerate tests/golden"


@pytest.mark.parametrize("tag,C,H,N,T", [("c64", 64, 4, 3, 10), ("c512", 512, 8, 2, 24)])
def test_attention_self(golden_dir, tag, C, H, N, T):
    case = f"attn_self_{tag}"
    g = load(golden_dir, case)
    P = GS.make_params(case, GS.attention_shapes(C))
    q, k, graph = GS.attention_case(case, C, N, T, T, self_att=True)
    check_checksum(g, {**P, "q": q, "graph": graph})
    P = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    x = q.clone().requires_grad_(True)
    y, att = O.attention(x, x, x, graph, P, H)
    assert O.rel_err(y, t(g["y"])) < FP32_TOL
    assert O.rel_err(att, t(g["att"])) < FP32_TOL
    # structural facts of the reference: zero graph row -> zero attention row
    assert float(att.view(H, N, T, T)[:, 0, 1].abs().sum()) == 0.0
    assert float(att.view(H, N, T, T)[:, 0, 2].abs().sum()) == 0.0
    (y * GS.randn(f"{case}/dy", *y.shape)).sum().backward()
    assert O.rel_err(x.grad, t(g["dx"])) < 1e-5
    GS.compare_grads(case, {k_: v.grad for k_, v in P.items()}, g, 1e-5)


@pytest.mark.parametrize("tag,C,H,N,T", [("c64", 64, 4, 3, 10), ("c512", 512, 8, 2, 24)])
@pytest.mark.parametrize("tq", [1, 3])
def test_attention_cross(golden_dir, tag, C, H, N, T, tq):
    case = f"attn_cross{tq}_{tag}"
    g = load(golden_dir, case)
    P = GS.make_params(case, GS.attention_shapes(C))
    q, k, graph = GS.attention_case(case, C, N, tq, T, self_att=False)
    check_checksum(g, {**P, "q": q, "k": k, "graph": graph})
    P = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    qq, kk = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
    y, att = O.attention(qq, kk, kk, graph, P, H)
    assert O.rel_err(y, t(g["y"])) < FP32_TOL
    assert O.rel_err(att, t(g["att"])) < FP32_TOL
    (y * GS.randn(f"{case}/dy", *y.shape)).sum().backward()
    assert O.rel_err(qq.grad, t(g["dq"])) < 1e-5
    assert O.rel_err(kk.grad, t(g["dk"])) < 1e-5
    GS.compare_grads(case, {k_: v.grad for k_, v in P.items()}, g, 1e-5)


@pytest.mark.parametrize("tag,C,H,N", [("c64", 64, 4, 3), ("c512", 512, 8, 2)])
def test_mha_causal(golden_dir, tag, C, H, N):
    case = f"mha_causal_{tag}"
    g = load(golden_dir, case)
    P = GS.make_params(case, GS.attention_shapes(C))
    q, _, _ = GS.attention_case(case, C, N, 6, 6, self_att=True)
    check_checksum(g, {**P, "q": q})
    P = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    x = q.clone().requires_grad_(True)
    y, _ = O.attention(x, x, x, None, P, H, causality=True, renorm="none")
    assert O.rel_err(y, t(g["y"])) < FP32_TOL
    (y * GS.randn(f"{case}/dy", *y.shape)).sum().backward()
    assert O.rel_err(x.grad, t(g["dx"])) < 1e-5
    GS.compare_grads(case, {k_: v.grad for k_, v in P.items()}, g, 1e-5)


@pytest.mark.parametrize("tag,C,H,N", [("c64", 64, 4, 5), ("c512", 512, 8, 3)])
def test_mha_single_token(golden_dir, tag, C, H, N):
    case = f"mha_token_{tag}"  # one token attending to itself: the decoder's self-attention (AttModel_x3.py:148)
    g = load(golden_dir, case)
    P = GS.make_params(case, GS.attention_shapes(C))
    q, _, _ = GS.attention_case(case, C, N, 1, 1, self_att=True)
    check_checksum(g, {**P, "q": q})
    P = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    x = q.clone().requires_grad_(True)
    y, _ = O.attention(x, x, x, None, P, H, causality=True, renorm="none")
    assert O.rel_err(y, t(g["y"])) < FP32_TOL
    (y * GS.randn(f"{case}/dy", *y.shape)).sum().backward()
    assert O.rel_err(x.grad, t(g["dx"])) < 1e-5
    GS.compare_grads(case, {k_: v.grad for k_, v in P.items()}, g, 1e-5)


@pytest.mark.parametrize("tag,C,H,N,T", [("c64", 64, 4, 3, 10), ("c512", 512, 8, 2, 24)])
def test_attention_graphmask(golden_dir, tag, C, H, N, T):
    case = f"attn_graphmask_{tag}"
    g = load(golden_dir, case)
    P = GS.make_params(case, GS.attention_shapes(C))
    q, _, graph = GS.attention_case(case, C, N, T, T, self_att=True)
    check_checksum(g, {**P, "q": q, "graph": graph})
    y, att = O.attention(q, q, q, graph, P, H, renorm="addeps")
    assert O.rel_err(y, t(g["y"])) < FP32_TOL
    assert O.rel_err(att, t(g["att"])) < FP32_TOL


@pytest.mark.parametrize("tag,C,N,T", [("c64", 64, 3, 10), ("c512", 512, 2, 24)])
def test_feedforward(golden_dir, tag, C, N, T):
    case = f"ffn_{tag}"
    g = load(golden_dir, case)
    P = GS.make_params(case, GS.feedforward_shapes(C))
    xin = GS.randn(f"{case}/x", N, T, C)
    check_checksum(g, {**P, "x": xin})
    P = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    x = xin.clone().requires_grad_(True)
    y = O.feedforward(x, P)
    assert O.rel_err(y, t(g["y"])) < FP32_TOL
    (y * GS.randn(f"{case}/dy", *y.shape)).sum().backward()
    assert O.rel_err(x.grad, t(g["dx"])) < 1e-5
    GS.compare_grads(case, {k_: v.grad for k_, v in P.items()}, g, 1e-5)


def test_layernorm(golden_dir):
    case = "layernorm"
    g = load(golden_dir, case)
    C = 512
    xin = GS.randn(f"{case}/x", 5, 7, C, scale=2.0)
    xin[0, 0, :] = 1.25
    gamma = GS.rand(f"{case}/gamma", C, lo=0.8, hi=1.2).requires_grad_(True)
    beta = GS.randn(f"{case}/beta", C, scale=0.1).requires_grad_(True)
    check_checksum(g, {"gamma": gamma.detach(), "beta": beta.detach(), "x": xin})
    x = xin.clone().requires_grad_(True)
    y = O.layer_norm(x, gamma, beta)
    yr = t(g["y"])
    assert O.rel_err(y, yr) < FP32_TOL
    # sigma == 0 row: output is exactly beta
    assert torch.equal(y[0, 0].detach(), beta.detach())
    (y * GS.randn(f"{case}/dy", *y.shape)).sum().backward()
    # the sigma=0 row's dx is (g - mean g)/1e-8: huge but finite; compare the regular rows tightly
    dx, dxr = x.grad, t(g["dx"])
    assert O.rel_err(dx[1:], dxr[1:]) < 1e-5
    assert torch.isfinite(dxr[0, 0]).all() == torch.isfinite(dx[0, 0]).all()
    assert O.rel_err(beta.grad, t(g["dbeta"])) < 1e-5


@pytest.mark.parametrize("zp", [True, False])
@pytest.mark.parametrize("sc", [True, False])
def test_embedding(golden_dir, zp, sc):
    case = f"embedding_zp{int(zp)}_sc{int(sc)}"
    g = load(golden_dir, case)
    table = GS.randn(f"{case}/table", 11, 64, scale=0.3)
    idx = GS.randint(f"{case}/idx", 0, 11, 4, 6)
    idx[0, 0], idx[0, 1] = 0, 10
    y = O.embedding_lookup(idx, table, sc)
    assert torch.equal(y, t(g["y"]))  # bit exact
    # gradient hole: row 0 (zeros_pad) or the LAST row (padding_idx=-1) receives no gradient (modules.py:34-41)
    dtab = t(g["dtable"])
    hole = 0 if zp else 10
    assert float(dtab[hole].abs().sum()) == 0.0


@pytest.mark.parametrize("kind", ["vis", "syb"])
def test_branch(golden_dir, kind):
    S = GS.SMALL
    case = f"branch_{kind}_c64"
    g = load(golden_dir, case)
    shapes = GS.branch_shapes(kind, S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"])
    P = GS.make_params(case, shapes)
    check_checksum(g, P)
    nfirst = S["V"] if kind == "vis" else S["M"]
    b = GS.branch_case(case, kind, S["B"], nfirst, S["Q"])
    P = {k_: v.clone().requires_grad_(True) for k_, v in P.items()}
    first = b["first"].clone().requires_grad_(True)
    taps = {}
    dec = O.branch_forward(P, kind, first, b["first_mask"], b["first_graph"], b["q_ipt"], b["q_graph"], b["q_mask"], True,
                           S["blocks"], S["heads"], taps=taps)
    # mask construction: bit exact, including the graph_cross == graph aliasing (blocks 2..5 share one mask)
    for i in range(S["blocks"]):
        expect = taps["graph_diag"] if i < 2 else taps["graph"]
        assert torch.equal(expect, t(g[f"tap/enc_graph_{i}"])), i
    assert torch.equal(taps["dec_mask"], t(g["tap/dec_mask"]))
    for i in range(S["blocks"]):
        assert O.rel_err(taps[f"enc_att_{i}"], t(g[f"tap/enc_att_{i}"])) < 1e-5, i
        assert O.rel_err(taps[f"enc_ffn_{i}"], t(g[f"tap/enc_ffn_{i}"])) < 1e-5, i
        assert O.rel_err(taps[f"dec_{i}"], t(g[f"tap/dec_{i}"])) < 1e-5, i
    assert O.rel_err(dec, t(g["dec"])) < 1e-5
    (dec * GS.randn(f"{case}/ddec", *dec.shape)).sum().backward()
    assert O.rel_err(first.grad, t(g["dfirst"])) < 1e-4
    grads = {k_: (v.grad if v.grad is not None else torch.zeros_like(v)) for k_, v in P.items()}
    GS.compare_grads(case, grads, g, 1e-4)


def test_full_heads_and_loss(golden_dir):
    S = GS.SMALL
    case = "full_c64"
    g = load(golden_dir, case)
    Pv = GS.make_params("branch_vis_c64", GS.branch_shapes("vis", S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]))
    Ps = GS.make_params("branch_syb_c64", GS.branch_shapes("syb", S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]))
    Ph = GS.make_params(case, GS.head_shapes(S["C"], S["ncls"]))
    check_checksum(g, Ph)
    params = {**{"att_vis_grid." + k: v for k, v in Pv.items()}, **{"att_syb." + k: v for k, v in Ps.items()}, **Ph}
    bv = GS.branch_case("branch_vis_c64", "vis", S["B"], S["V"], S["Q"])
    bs = GS.branch_case("branch_syb_c64", "syb", S["B"], S["M"], S["Q"])
    batch = dict(vis_fea=bv["first"], vis_fea_mask=bv["first_mask"], q_ipt=bv["q_ipt"], q_ipt_graph=bv["q_graph"],
                 q_ipt_mask=bv["q_mask"], syb_ipt=t(g["syb_ipt"]), macro_node_mask=bs["first_mask"],
                 macro_graph_ipt=bs["first_graph"], answer=GS.randint(f"{case}/answer", 0, S["ncls"], S["B"]))
    loss, (lc, lv, ls), _, _ = O.encoder_step(params, batch, S["blocks"], S["heads"])
    assert O.rel_err(lc, t(g["logits_concat"])) < 1e-5
    assert O.rel_err(lv, t(g["logits_vis"])) < 1e-5
    assert O.rel_err(ls, t(g["logits_syb"])) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))


def test_state_dict_contract(golden_dir):
    """The parameter names/shapes the drop-in must expose were dumped from the live reference."""
    S = GS.SMALL
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys_c64.json")))
    ref = {k: tuple(s) for k, s in keys}
    for pref, kind in (("att_vis_grid.", "vis"), ("att_syb.", "syb")):
        for k, s in GS.branch_shapes(kind, S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]):
            assert ref[pref + k] == tuple(s), k
    for k, s in GS.head_shapes(S["C"], S["ncls"]):
        assert ref[k] == tuple(s), k


def test_g_weighted_softmax_identity():
    """softmax -> *G -> L1 renorm == G-weighted softmax (what the kernels compute)."""
    torch.manual_seed(1)
    S = torch.randn(4, 7, 9, dtype=torch.float64)
    G = (torch.rand(4, 7, 9, dtype=torch.float64) < 0.4).double()
    G[0, 0] = 0
    P = torch.softmax(S, -1)
    A = P * G
    W_ref = A / A.abs().sum(-1, keepdim=True).clamp_min(1e-12)
    e = torch.exp(S - S.max(-1, keepdim=True).values) * G
    W = e / e.sum(-1, keepdim=True).clamp_min(1e-300)
    W[0, 0] = 0
    assert torch.allclose(W, W_ref, atol=1e-14)


@pytest.mark.parametrize("case", sorted(GS.MIL_CASES))
def test_mil_nce_oracle_vs_live_reference_golden(golden_dir, case):
    """O.mil_nce (AttModel_x3.py:339-379, 441) against the live reference's MIL_NCE: output, mil_nce_obj and the gradients of every
    parameter the only_obj forward uses (marco_mlp and the macro-node word rows get none: the `.detach()` at :354)."""
    c = GS.MIL_CASES[case]
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    P0 = GS.make_params(case, GS.mil_nce_shapes(c["h"]))
    assert abs(GS.checksum({**P0, "vis": GS.mil_nce_case(case, c["B"], c["V"], c["M"], c["topN"])["vis_fea"]}) - float(g["checksum"])) < 1e-6 * abs(float(g["checksum"]))
    P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
    b = GS.mil_nce_case(case, c["B"], c["V"], c["M"], c["topN"])
    out, obj = O.mil_nce(P, b["vis_fea"], b["macro_ipt"], b["macro_obj_loc"], b["pos"], b["neg"], b["mask"])
    assert O.rel_err(out, torch.from_numpy(g["out"])) < 2e-6
    assert abs(float(obj) - float(g["obj"])) < 2e-6 * max(1.0, abs(float(g["obj"])))
    w = GS.randn(f"{case}/dout", *out.shape)
    ((out * w).sum() + 3.0 * obj).backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()}
    GS.compare_grads(case, grads, g, 5e-6)
    assert float(grads["marco_mlp.0.weight"].abs().sum()) == 0.0
