"""Generate tests/golden/*.npz by running the LIVE, UNMODIFIED reference (/root/reference/models/modules.py
and AttModel_x3.py) on the deterministic cases of oracle/golden_spec.py.  TEST INFRASTRUCTURE ONLY.

Run in the authoring container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py [--ref /root/reference]

Harness -- no reference file is edited or copied:
  * sys.path gets <ref>/models because AttModel_x3.py does `from modules import *` (AttModel_x3.py:9);
  * on a CPU-only box Tensor.cuda / Module.cuda are patched to identity (hard-coded .cuda() at
    modules.py:168,174,179,261,269,273 and AttModel_x3.py:101,104-108,141,146);
  * models are built under torch.no_grad() (in-place write into a leaf Parameter, AttModel_x3.py:38,170);
  * dropout_rate = 0.0; LN affine randomised (golden_spec.make_params).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import golden_spec as GS  # noqa: E402


def import_reference(ref_root: str):
    sys.path.insert(0, os.path.join(ref_root, "models"))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import warnings
    warnings.filterwarnings("ignore")
    import modules as M  # type: ignore
    import AttModel_x3 as A  # type: ignore
    return M, A


def save(path, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.1f} KiB")


def load_params(module, params):
    sd = module.state_dict()
    assert set(sd.keys()) == set(params.keys()), (sorted(set(sd) ^ set(params)))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(params[k].shape), (k, sd[k].shape, params[k].shape)
    module.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)


def grads_of(module, keys):
    named = dict(module.named_parameters())
    return {k: (named[k].grad.detach().clone() if named[k].grad is not None else torch.zeros_like(named[k])) for k in keys}


def attention_goldens(M, out_dir):
    for tag, C, H in (("c64", 64, 4), ("c512", GS.WIDE["C"], GS.WIDE["heads"])):
        N, T = (3, 10) if C == 64 else (2, 24)
        # --- new_multihead_attention, self attention -----------------------------------------
        case = f"attn_self_{tag}"
        P = GS.make_params(case, GS.attention_shapes(C))
        q, k, graph = GS.attention_case(case, C, N, T, T, self_att=True)
        m = M.new_multihead_attention(C, H, return_att=True)
        load_params(m, P)
        x = q.clone().requires_grad_(True)
        y, att = m(x, x, x, graph)
        w = GS.randn(f"{case}/dy", *y.shape)
        (y * w).sum().backward()
        g = grads_of(m, list(P.keys()))
        save(os.path.join(out_dir, f"{case}.npz"), y=y, att=att, dx=x.grad, checksum=GS.checksum({**P, "q": q, "graph": graph}),
             **GS.pack_grads(case, g))

        # --- new_multihead_attention, cross attention (decoder style: Tq=1 and Tq=3) ---------
        for tq in (1, 3):
            case = f"attn_cross{tq}_{tag}"
            P = GS.make_params(case, GS.attention_shapes(C))
            q, k, graph = GS.attention_case(case, C, N, tq, T, self_att=False)
            m = M.new_multihead_attention(C, H, return_att=True)
            load_params(m, P)
            qq = q.clone().requires_grad_(True)
            kk = k.clone().requires_grad_(True)
            y, att = m(qq, kk, kk, graph)
            w = GS.randn(f"{case}/dy", *y.shape)
            (y * w).sum().backward()
            g = grads_of(m, list(P.keys()))
            save(os.path.join(out_dir, f"{case}.npz"), y=y, att=att, dq=qq.grad, dk=kk.grad,
                 checksum=GS.checksum({**P, "q": q, "k": k, "graph": graph}), **GS.pack_grads(case, g))

        # --- multihead_attention (causal) ------------------------------------------------------
        case = f"mha_causal_{tag}"
        P = GS.make_params(case, GS.attention_shapes(C))
        q, k, _ = GS.attention_case(case, C, N, 6, 6, self_att=True)
        m = M.multihead_attention(C, H, causality=True)
        load_params(m, P)
        x = q.clone().requires_grad_(True)
        y = m(x, x, x)
        w = GS.randn(f"{case}/dy", *y.shape)
        (y * w).sum().backward()
        g = grads_of(m, list(P.keys()))
        save(os.path.join(out_dir, f"{case}.npz"), y=y, dx=x.grad, checksum=GS.checksum({**P, "q": q}),
             **GS.pack_grads(case, g))

        # --- new_multihead_attention_with_graph_mask -------------------------------------------
        case = f"attn_graphmask_{tag}"
        P = GS.make_params(case, GS.attention_shapes(C))
        q, k, graph = GS.attention_case(case, C, N, T, T, self_att=True)
        m = M.new_multihead_attention_with_graph_mask(C, H, return_att=True)
        load_params(m, P)
        y, att = m(q, q, q, None, graph)
        save(os.path.join(out_dir, f"{case}.npz"), y=y, att=att, checksum=GS.checksum({**P, "q": q, "graph": graph}))

        # --- feedforward ------------------------------------------------------------------------
        case = f"ffn_{tag}"
        P = GS.make_params(case, GS.feedforward_shapes(C))
        xin = GS.randn(f"{case}/x", N, T, C)
        m = M.feedforward(C, [4 * C, C])
        load_params(m, P)
        x = xin.clone().requires_grad_(True)
        y = m(x)
        w = GS.randn(f"{case}/dy", *y.shape)
        (y * w).sum().backward()
        g = grads_of(m, list(P.keys()))
        save(os.path.join(out_dir, f"{case}.npz"), y=y, dx=x.grad, checksum=GS.checksum({**P, "x": xin}),
             **GS.pack_grads(case, g))

    # --- layer_normalization incl. a constant (sigma = 0) row ---------------------------------------
    case = "layernorm"
    C = 512
    xin = GS.randn(f"{case}/x", 5, 7, C, scale=2.0)
    xin[0, 0, :] = 1.25
    P = {"gamma": GS.rand(f"{case}/gamma", C, lo=0.8, hi=1.2), "beta": GS.randn(f"{case}/beta", C, scale=0.1)}
    m = M.layer_normalization(C)
    load_params(m, P)
    x = xin.clone().requires_grad_(True)
    y = m(x)
    w = GS.randn(f"{case}/dy", *y.shape)
    (y * w).sum().backward()
    save(os.path.join(out_dir, f"{case}.npz"), y=y, dx=x.grad, dgamma=m.gamma.grad, dbeta=m.beta.grad,
         checksum=GS.checksum({**P, "x": xin}))

    # --- embedding (zeros_pad / scale combinations, gradient holes) ------------------------------------
    for zp in (True, False):
        for sc in (True, False):
            case = f"embedding_zp{int(zp)}_sc{int(sc)}"
            table = GS.randn(f"{case}/table", 11, 64, scale=0.3)
            idx = GS.randint(f"{case}/idx", 0, 11, 4, 6)
            idx[0, 0], idx[0, 1] = 0, 10
            m = M.embedding(11, 64, zeros_pad=zp, scale=sc)
            with torch.no_grad():
                m.lookup_table.copy_(table)
            y = m(idx)
            w = GS.randn(f"{case}/dy", *y.shape)
            (y * w).sum().backward()
            save(os.path.join(out_dir, f"{case}.npz"), y=y, dtable=m.lookup_table.grad, checksum=GS.checksum({"t": table}))


def branch_goldens(M, A, out_dir):
    S = GS.SMALL
    glove = types.SimpleNamespace(vectors=torch.zeros(4, GS.E_GLOVE))
    for kind in ("vis", "syb"):
        case = f"branch_{kind}_c64"
        shapes = GS.branch_shapes(kind, S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"])
        P = GS.make_params(case, shapes)
        with torch.no_grad():
            if kind == "vis":
                m = A.AttModel_vis_grid(glove, S["C"], S["maxlen"], S["maxlen_q"], S["blocks"], S["heads"], 0.0, S["maxlen_v"], S["ncls"])
            else:
                m = A.AttModel_syb(glove, S["C"], S["maxlen"], S["maxlen_q"], S["blocks"], S["heads"], 0.0, S["ncls"])
        # parameter-name contract: the live reference must agree with golden_spec.branch_shapes
        ref_shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert ref_shapes == shapes, [a for a, b in zip(ref_shapes, shapes) if a != b][:5]
        full = {k: v.clone() for k, v in P.items()}
        big = m.state_dict()["syb_emb.weight"].clone()
        big[: GS.SMALL_VOCAB] = P["syb_emb.weight"]
        full["syb_emb.weight"] = big
        load_params(m, full)

        nfirst = S["V"] if kind == "vis" else S["M"]
        b = GS.branch_case(case, kind, S["B"], nfirst, S["Q"])
        taps = {}

        def hook(name):
            def f(mod, inp, out):
                taps[name] = (out[0] if isinstance(out, tuple) else out).detach().clone()
            return f

        def pre(name):
            def f(mod, inp):
                taps[name] = inp[3].detach().clone()
            return f

        for i in range(S["blocks"]):
            getattr(m, f"enc_self_attention_{i}").register_forward_hook(hook(f"enc_att_{i}"))
            getattr(m, f"enc_feed_forward_{i}").register_forward_hook(hook(f"enc_ffn_{i}"))
            getattr(m, f"dec_feed_forward_{i}").register_forward_hook(hook(f"dec_{i}"))
            getattr(m, f"enc_self_attention_{i}").register_forward_pre_hook(pre(f"enc_graph_{i}"))
        getattr(m, "dec_vanilla_attention_0").register_forward_pre_hook(pre("dec_mask"))

        first = b["first"].clone().requires_grad_(True)
        if kind == "vis":
            dec = m(first, b["first_mask"], b["q_ipt"], b["q_graph"], b["q_mask"], True)
        else:
            dec = m(first, b["first_mask"], b["first_graph"], b["q_ipt"], b["q_graph"], b["q_mask"], True)
        w = GS.randn(f"{case}/ddec", *dec.shape)
        (dec * w).sum().backward()
        named = dict(m.named_parameters())
        grads = {}
        for k in P:
            g = named[k].grad
            g = torch.zeros_like(named[k]) if g is None else g
            if k == "syb_emb.weight":
                assert float(g[GS.SMALL_VOCAB:].abs().sum()) == 0.0
                g = g[: GS.SMALL_VOCAB]
            grads[k] = g
        save(os.path.join(out_dir, f"{case}.npz"), dec=dec, dfirst=first.grad, checksum=GS.checksum(P),
             **{f"tap/{k}": v for k, v in taps.items()}, **GS.pack_grads(case, grads))

    # --- heads + loss through the full reference AttModel (MIL_NCE runs as the reference; its output
    #     `new_macro_ipt` is captured and handed to the oracle as `syb_ipt`) --------------------------
    case = "full_c64"
    hs = 16  # hidden_size_mil
    with torch.no_grad():
        full_model = A.AttModel(glove, S["C"], hs, S["ncls"], S["maxlen_q"], S["maxlen"], S["maxlen_v"], S["blocks"], S["heads"],
                                0.0, 0.1, 5, True)
    sd = full_model.state_dict()
    keys = [(k, tuple(v.shape)) for k, v in sd.items()]
    with open(os.path.join(out_dir, "state_dict_keys_c64.json"), "w") as f:
        json.dump(keys, f, indent=0)
    Pv = GS.make_params("branch_vis_c64", GS.branch_shapes("vis", S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]))
    Ps = GS.make_params("branch_syb_c64", GS.branch_shapes("syb", S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]))
    Ph = GS.make_params(case, GS.head_shapes(S["C"], S["ncls"]))
    new_sd = {k: v.clone() for k, v in sd.items()}
    for pref, P in (("att_vis_grid.", Pv), ("att_syb.", Ps)):
        for k, v in P.items():
            if k == "syb_emb.weight":
                new_sd[pref + k][: GS.SMALL_VOCAB] = v
            else:
                new_sd[pref + k] = v.clone()
    for k, v in Ph.items():
        new_sd[k] = v.clone()
    new_sd["MIL_NCE.syb_emb.weight"][: GS.SMALL_VOCAB] = GS.randn(f"{case}/mil_table", GS.SMALL_VOCAB, GS.E_GLOVE, scale=0.5)
    full_model.load_state_dict(new_sd)
    full_model.eval()

    bv = GS.branch_case("branch_vis_c64", "vis", S["B"], S["V"], S["Q"])
    bs = GS.branch_case("branch_syb_c64", "syb", S["B"], S["M"], S["Q"])
    B, V, Mn = S["B"], S["V"], S["M"]
    macro_ipt = GS.randint(f"{case}/macro_ipt", 0, GS.SMALL_VOCAB - 1, B, Mn)
    macro_obj_loc = torch.full((B, V), -1, dtype=torch.int64)
    for b in range(B):
        macro_obj_loc[b, :3] = torch.tensor([0, 2, 4])
    topn = 2
    pos = GS.randint(f"{case}/pos", 0, GS.SMALL_VOCAB - 1, B, V, topn)
    neg = GS.randint(f"{case}/neg", 0, GS.SMALL_VOCAB - 1, B, V, topn)
    omask = torch.ones(B, V, topn, dtype=torch.int32)
    e = torch.empty((B, 0))
    answer = GS.randint(f"{case}/answer", 0, S["ncls"], B)
    cap = {}
    full_model.MIL_NCE.register_forward_hook(lambda mod, inp, out: cap.__setitem__("syb_ipt", out[0].detach().clone()))
    with torch.no_grad():
        lc, lv, ls, mil_obj, mil_rel = full_model(bv["first"], bv["first_mask"], bv["q_ipt"], bv["q_mask"], bv["q_graph"],
                                                  macro_ipt, bs["first_mask"], bs["first_graph"], macro_obj_loc, pos, neg, omask,
                                                  e, e, e, e, decMask=True, mcb=False)
        lsm = (torch.log_softmax(lv, -1) + torch.log_softmax(ls, -1) + torch.log_softmax(lc, -1)) / 3
        one_hot = torch.zeros_like(lc)
        one_hot.scatter_(1, answer.view(-1, 1), 1)
        one_hot = M.label_smoothing()(one_hot)
        loss = (-(one_hot * lsm).sum(-1)).mean()
    save(os.path.join(out_dir, f"{case}.npz"), syb_ipt=cap["syb_ipt"], logits_concat=lc, logits_vis=lv, logits_syb=ls, loss=loss,
         checksum=GS.checksum(Ph))


def token_goldens(M, out_dir):
    """multihead_attention (causal) on ONE token per sample attending to itself -- the decoder's self-attention
    (AttModel_x3.py:148): the degenerate case savqa_b200 runs as a single N = C GEMM (functional.TokenSelfAttentionFn)."""
    for tag, C, H in (("c64", 64, 4), ("c512", GS.WIDE["C"], GS.WIDE["heads"])):
        N = 5 if C == 64 else 3
        case = f"mha_token_{tag}"
        P = GS.make_params(case, GS.attention_shapes(C))
        q, _, _ = GS.attention_case(case, C, N, 1, 1, self_att=True)
        m = M.multihead_attention(C, H, causality=True)
        load_params(m, P)
        x = q.clone().requires_grad_(True)
        y = m(x, x, x)
        w = GS.randn(f"{case}/dy", *y.shape)
        (y * w).sum().backward()
        g = grads_of(m, list(P.keys()))
        save(os.path.join(out_dir, f"{case}.npz"), y=y, dx=x.grad, checksum=GS.checksum({**P, "q": q}), **GS.pack_grads(case, g))


def mil_nce_goldens(A, out_dir):
    """The live reference's MIL_NCE (AttModel_x3.py:285-443, only_obj=True): outputs and gradients of every parameter it uses
    for L = sum(out * w) + 3 * mil_nce_obj."""
    glove = types.SimpleNamespace(vectors=torch.zeros(4, GS.E_GLOVE))
    for case, c in GS.MIL_CASES.items():
        P = GS.make_params(case, GS.mil_nce_shapes(c["h"]))
        with torch.no_grad():
            m = A.MIL_NCE(glove, c["h"], 0.0, 5, True)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        for k, v in P.items():
            if k == "syb_emb.weight":
                sd[k][: GS.SMALL_VOCAB] = v
            else:
                assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
                sd[k] = v.clone()
        m.load_state_dict(sd)
        b = GS.mil_nce_case(case, c["B"], c["V"], c["M"], c["topN"])
        e = torch.empty((c["B"], 0))
        out, obj, rel = m(b["vis_fea"], b["macro_ipt"], b["macro_obj_loc"], b["pos"], b["neg"], b["mask"], e, e, e, e)
        assert rel == 0
        w = GS.randn(f"{case}/dout", *out.shape)
        ((out * w).sum() + 3.0 * obj).backward()
        named = dict(m.named_parameters())
        grads = {}
        for k in P:
            g = named[k].grad
            g = torch.zeros_like(named[k]) if g is None else g
            if k == "syb_emb.weight":
                assert float(g[GS.SMALL_VOCAB:].abs().sum()) == 0.0
                g = g[: GS.SMALL_VOCAB]
            grads[k] = g
        save(os.path.join(out_dir, f"{case}.npz"), out=out, obj=obj, checksum=GS.checksum({**P, "vis": b["vis_fea"]}),
             **GS.pack_grads(case, grads))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default="", help="'token' / 'mil': (re)generate only the one-token self-attention / the MIL_NCE fixtures")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)  # bit-stable reductions while generating
    M, A = import_reference(args.ref)
    if args.only == "mil":
        mil_nce_goldens(A, args.out)
    else:
        if args.only != "token":
            attention_goldens(M, args.out)
            branch_goldens(M, A, args.out)
            mil_nce_goldens(A, args.out)
        token_goldens(M, args.out)
    with open(os.path.join(args.out, "MANIFEST.json"), "w") as f:
        json.dump({"torch": torch.__version__, "generator": "oracle/make_golden.py", "reference": "Peixixiong/Structured-Alignment-VQA",
                   "files": sorted(x for x in os.listdir(args.out) if x.endswith(".npz"))}, f, indent=1)


if __name__ == "__main__":
    main()
