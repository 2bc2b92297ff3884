"""Parity cases shared by the CPU host-logic tests (kernels replaced by tests/fake_ops.py) and the GPU parity tests
(real kernels through the C ABI).  Each case builds the seeded inputs of oracle/golden_spec.py, runs OUR module on
`device`, and checks outputs and gradients against (i) the golden vectors produced by the live reference and
(ii) the oracle, with per-tensor tolerances calibrated by the oracle's bf16-operand emulation (parity_util.check).
"""
import types

import torch

from oracle import golden_spec as GS
from oracle import savqa_oracle as O
from parity_util import check, grads_of, load, set_params, t

BF = torch.bfloat16


def _oracle_attention(P0, q, k, graph, H, dy, operand_dtype, self_att, **kw):
    P = {k_: v.clone().requires_grad_(True) for k_, v in P0.items()}
    qq = q.clone().requires_grad_(True)
    kk = qq if self_att else k.clone().requires_grad_(True)
    y, att = O.attention(qq, kk, kk, graph, P, H, operand_dtype=operand_dtype, **kw)
    (y * dy).sum().backward()
    out = {"y": y.detach(), "att": att.detach(), "dq": qq.grad, "dk": kk.grad}
    out.update({k_: (v.grad if v.grad is not None else torch.zeros_like(v)) for k_, v in P.items()})
    return out


def attention_case(M, golden_dir, device, case, C, H, N, Tq, Tk, self_att, kind="new"):
    """kind: "new" = new_multihead_attention, "mha" = multihead_attention(causal), "gm" = ..._with_graph_mask."""
    g = load(golden_dir, case)
    P0 = GS.make_params(case, GS.attention_shapes(C))
    q, k, graph = GS.attention_case(case, C, N, Tq, Tk, self_att=self_att)
    dy = GS.randn(f"{case}/dy", N, Tq, C)
    kw = {"new": {}, "mha": dict(causality=True, renorm="none"), "gm": dict(renorm="addeps")}[kind]
    ref = _oracle_attention(P0, q, k, None if kind == "mha" else graph, H, dy, None, self_att, **kw)
    emu = _oracle_attention(P0, q, k, None if kind == "mha" else graph, H, dy, BF, self_att, **kw)
    # the oracle itself is pinned to the live reference's golden output
    assert O.rel_err(ref["y"], t(g["y"])) < 2e-6

    if kind == "new":
        m = M.new_multihead_attention(C, H, return_att=True)
    elif kind == "mha":
        m = M.multihead_attention(C, H, causality=True)
    else:
        m = M.new_multihead_attention_with_graph_mask(C, H, return_att=True)
    set_params(m, P0)
    m = m.to(device)
    qq = q.clone().to(device).requires_grad_(True)
    kk = qq if self_att else k.clone().to(device).requires_grad_(True)
    gd = graph.to(device)
    if kind == "new":
        y, att = m(qq, kk, kk, gd)
    elif kind == "mha":
        y, att = m(qq, kk, kk), None
    else:
        y, att = m(qq, kk, kk, None, gd)
    assert y.shape == (N, Tq, C)
    errs = {"y": check(f"{case}: output", y, t(g["y"]), emu["y"])}
    if att is not None:
        assert att.shape == (H * N, Tq, Tk)
        errs["att"] = check(f"{case}: attention probabilities", att, t(g["att"]), emu["att"])
        if Tq >= 3 and Tk >= 4 and kind == "new":
            a4 = att.detach().cpu().view(H, N, Tq, Tk)
            # a query row without any graph edge, and one whose only edge is key-masked: exactly zero rows
            assert float(a4[:, 0, 1].abs().sum()) == 0.0 and float(a4[:, 0, 2].abs().sum()) == 0.0
    (y * dy.to(device)).sum().backward()
    # rows whose LayerNorm input is constant (zero query row) carry a 1/eps = 1e8 scaled gradient: compare apart
    zero_q = (q.abs().sum(-1) == 0)
    errs["dq"] = check(f"{case}: d queries", qq.grad, ref["dq"], emu["dq"], mask=zero_q)
    if not self_att:
        errs["dk"] = check(f"{case}: d keys", kk.grad, ref["dk"], emu["dk"])
    ours = grads_of(m, P0.keys())
    for k_ in P0:
        errs[k_] = check(f"{case}: grad {k_}", ours[k_], ref[k_], emu[k_])
    return errs


def feedforward_case(M, golden_dir, device, case, C, N, T):
    g = load(golden_dir, case)
    P0 = GS.make_params(case, GS.feedforward_shapes(C))
    xin = GS.randn(f"{case}/x", N, T, C)
    dy = GS.randn(f"{case}/dy", N, T, C)

    def run(od):
        P = {k_: v.clone().requires_grad_(True) for k_, v in P0.items()}
        x = xin.clone().requires_grad_(True)
        y = O.feedforward(x, P, od)
        (y * dy).sum().backward()
        return dict(y=y.detach(), dx=x.grad, **{k_: v.grad for k_, v in P.items()})

    ref, emu = run(None), run(BF)
    assert O.rel_err(ref["y"], t(g["y"])) < 2e-6
    m = M.feedforward(C, [4 * C, C])
    set_params(m, P0)
    m = m.to(device)
    x = xin.clone().to(device).requires_grad_(True)
    y = m(x)
    errs = {"y": check(f"{case}: output", y, t(g["y"]), emu["y"])}
    (y * dy.to(device)).sum().backward()
    errs["dx"] = check(f"{case}: dx", x.grad, ref["dx"], emu["dx"])
    ours = grads_of(m, P0.keys())
    for k_ in P0:
        errs[k_] = check(f"{case}: grad {k_}", ours[k_], ref[k_], emu[k_])
    return errs


def layernorm_case(M, golden_dir, device):
    g = load(golden_dir, "layernorm")
    C = 512
    xin = GS.randn("layernorm/x", 5, 7, C, scale=2.0)
    xin[0, 0, :] = 1.25
    m = M.layer_normalization(C)
    set_params(m, {"gamma": GS.rand("layernorm/gamma", C, lo=0.8, hi=1.2), "beta": GS.randn("layernorm/beta", C, scale=0.1)})
    m = m.to(device)
    x = xin.clone().to(device).requires_grad_(True)
    y = m(x)
    check("layernorm: y", y, t(g["y"]), floor=2e-6)  # fp32 kernel
    assert torch.equal(y[0, 0].detach().cpu(), m.beta.detach().cpu())  # sigma == 0 row -> exactly beta
    (y * GS.randn("layernorm/dy", *y.shape).to(device)).sum().backward()
    dx, dxr = x.grad.detach().cpu(), t(g["dx"])
    check("layernorm: dx", dx[1:], dxr[1:], floor=1e-5)
    check("layernorm: dx (row 0, regular)", dx[0, 1:], dxr[0, 1:], floor=1e-5)
    check("layernorm: dx (sigma == 0 row)", dx[0, 0], dxr[0, 0], floor=1e-4)
    check("layernorm: dgamma", m.gamma.grad, t(g["dgamma"]), floor=1e-4)
    check("layernorm: dbeta", m.beta.grad, t(g["dbeta"]), floor=1e-5)


def embedding_cases(M, golden_dir, device):
    for zp in (True, False):
        for sc in (True, False):
            case = f"embedding_zp{int(zp)}_sc{int(sc)}"
            g = load(golden_dir, case)
            table = GS.randn(f"{case}/table", 11, 64, scale=0.3)
            idx = GS.randint(f"{case}/idx", 0, 11, 4, 6)
            idx[0, 0], idx[0, 1] = 0, 10
            e = M.embedding(11, 64, zeros_pad=zp, scale=sc)
            with torch.no_grad():
                e.lookup_table.copy_(table)
            e = e.to(device)
            y = e(idx.to(device))
            assert torch.equal(y.detach().cpu(), t(g["y"])), case  # gather: bit exact
            (y * GS.randn(f"{case}/dy", *y.shape).to(device)).sum().backward()
            check(f"{case}: dtable", e.lookup_table.grad, t(g["dtable"]), floor=1e-6)
            hole = 0 if zp else 10
            assert float(e.lookup_table.grad[hole].abs().sum()) == 0.0  # padding_idx gradient hole (modules.py:34-41)


def make_branch(A, kind, S, vocab_rows=None):
    glove = types.SimpleNamespace(vectors=torch.zeros(4, 300))
    saved = A.VOCAB_ROWS
    if vocab_rows is not None:
        A.VOCAB_ROWS = vocab_rows
    try:
        if kind == "vis":
            return A.AttModel_vis_grid(glove, S["C"], S["maxlen"], S["maxlen_q"], S["blocks"], S["heads"], 0.0, S["maxlen_v"], S["ncls"])
        return A.AttModel_syb(glove, S["C"], S["maxlen"], S["maxlen_q"], S["blocks"], S["heads"], 0.0, S["ncls"])
    finally:
        A.VOCAB_ROWS = saved


def branch_case(A, golden_dir, device, kind):
    S = GS.SMALL
    case = f"branch_{kind}_c64"
    g = load(golden_dir, case)
    shapes = GS.branch_shapes(kind, S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"])
    P0 = GS.make_params(case, shapes)
    nfirst = S["V"] if kind == "vis" else S["M"]
    b = GS.branch_case(case, kind, S["B"], nfirst, S["Q"])
    ddec = GS.randn(f"{case}/ddec", S["B"], 1, S["C"])

    def run(od):
        P = {k_: v.clone().requires_grad_(True) for k_, v in P0.items()}
        first = b["first"].clone().requires_grad_(True)
        taps = {}
        dec = O.branch_forward(P, kind, first, b["first_mask"], b["first_graph"], b["q_ipt"], b["q_graph"], b["q_mask"], True,
                               S["blocks"], S["heads"], operand_dtype=od, taps=taps)
        (dec * ddec).sum().backward()
        out = dict(dec=dec.detach(), dfirst=first.grad, taps=taps)
        out.update({k_: (v.grad if v.grad is not None else torch.zeros_like(v)) for k_, v in P.items()})
        return out

    ref, emu = run(None), run(BF)
    assert O.rel_err(ref["dec"], t(g["dec"])) < 1e-5
    m = make_branch(A, kind, S, vocab_rows=GS.SMALL_VOCAB)
    set_params(m, P0)  # strict=True: the key set equals the reference's
    m = m.to(device)
    first = b["first"].clone().to(device).requires_grad_(True)
    dv = lambda x: None if x is None else x.to(device)  # noqa: E731
    if kind == "vis":
        dec = m(first, dv(b["first_mask"]), dv(b["q_ipt"]), dv(b["q_graph"]), dv(b["q_mask"]), True)
    else:
        dec = m(first, dv(b["first_mask"]), dv(b["first_graph"]), dv(b["q_ipt"]), dv(b["q_graph"]), dv(b["q_mask"]), True)
    assert dec.shape == (S["B"], 1, S["C"])
    errs = {"dec": check(f"{case}: decoder output", dec, t(g["dec"]), emu["dec"])}
    (dec * ddec.to(device)).sum().backward()
    errs["dfirst"] = check(f"{case}: d first_ipt", first.grad, ref["dfirst"], emu["dfirst"], floor=2e-2, factor=6.0)
    ours = grads_of(m, P0.keys())
    for k_ in P0:
        if float(ref[k_].abs().sum()) == 0.0:  # parameters the reference's forward never uses
            assert float(ours[k_].abs().sum()) == 0.0, k_
            continue
        # 12 stacked bf16-operand blocks: the emulation itself sits at 1e-2..5e-2 on the K/Q projection gradients
        errs[k_] = check(f"{case}: grad {k_}", ours[k_], ref[k_], emu[k_], floor=2e-2, factor=6.0)
    return errs


def headline_step_case(device, batch_size=128, vocab_rows=5000, seed=11, floor=2e-2, factor=6.0, verbose=False):
    """The composition the headline number runs, against the oracle: GQA-shaped batch (T = 56 visual / 128 symbolic tokens per
    sample, C = 512, 8 heads, 6 + 6 blocks per branch), the REAL 12-block model (small vocabulary), train.EncoderTrainer's bound
    path (flat buffers, bf16 mirror, fused decoder K/V, side-stream weight gradients): loss, the three logit tensors, every
    dense parameter gradient of the flat buffer and the row-sparse word-table gradients vs O.encoder_step(...).backward() in
    CPU fp32 (main_itp_ddp_tar_super_node.py:321-345, 363), tolerances calibrated by the oracle's bf16-operand emulation."""
    from savqa_b200 import functional as Fn
    from savqa_b200 import ops, synthetic, train
    cfg = synthetic.GQA_SHAPED
    model = synthetic.build_model(cfg, vocab_rows=vocab_rows)
    batch = synthetic.make_batch(cfg, batch_size, seed=seed, vocab_rows=vocab_rows)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}

    def run(od):
        P = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
        loss, logits, _, _ = O.encoder_step(P, batch, cfg["blocks"], cfg["heads"], operand_dtype=od)
        loss.backward()
        return dict(loss=loss.detach(), logits=[l.detach() for l in logits],
                    grads={k: v.grad for k, v in P.items() if v.dtype.is_floating_point and v.grad is not None})

    ref, emu = run(None), run(BF)
    model = model.to(device)
    b = {k: v.to(device) for k, v in batch.items()}
    tr = train.EncoderTrainer(model, lr=1e-4, rowsparse=True)
    tr.prepare(b)
    tr.flat_grad.zero_()
    for t_ in tr.tables:
        t_._savqa_rowlog.clear()
        t_._savqa_rowlog.on_grad = None  # keep the (row id, row gradient) lists for the comparison below instead of applying them
    loss = tr._forward_backward(b)
    if str(device) != "cpu":
        Fn.join_wgrad_streams()
        torch.cuda.synchronize()
    errs = {"loss": abs(float(loss) - float(ref["loss"])) / abs(float(ref["loss"]))}
    assert errs["loss"] < 2e-3 + 3 * abs(float(emu["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])), errs
    with torch.no_grad():
        logits = model.encoder_step(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["syb_ipt"],
                                    b["macro_node_mask"], b["macro_graph_ipt"], True)
    for name, got, r, e in zip(("logits_concat", "logits_vis", "logits_syb"), logits, ref["logits"], emu["logits"]):
        errs[name] = check(f"headline: {name}", got, r, e, floor=1e-2, factor=3.0)  # <= 1e-2 end to end in bf16
    names = {id(p): k for k, p in model.named_parameters()}
    for p in tr.dense:
        k = names[id(p)]
        errs[k] = check(f"headline: grad {k}", p.grad, ref["grads"][k], emu["grads"][k], floor=floor, factor=factor)
    # parameters outside the flat buffers must be exactly the ones the reference gives no (or an all-zero) gradient
    dense = {names[id(p)] for p in tr.dense}
    tables = {"att_vis_grid.syb_emb.weight", "att_syb.syb_emb.weight"}
    for k, g_ in ref["grads"].items():
        if k not in dense and k not in tables:
            assert float(g_.abs().sum()) == 0.0, f"{k} has a reference gradient but is not in the trainer's flat buffers"
    # row-sparse word-table gradients: (row id, row gradient) lists -> dense, vs the reference's dense table gradient
    for t_, k in zip(tr.tables, ("att_vis_grid.syb_emb.weight", "att_syb.syb_emb.weight")):
        dense_g = torch.zeros_like(t_.weight.data)
        for idx, rows, scale, skip in t_._savqa_rowlog.pending:
            ops.scatter_add_rows(dense_g, idx, rows, scale=scale, skip_row=skip)
        errs[k] = check(f"headline: grad {k}", dense_g, ref["grads"][k], emu["grads"][k], floor=floor, factor=factor)
        t_._savqa_rowlog.clear()
    if verbose:
        worst = sorted(((v, k) for k, v in errs.items()), reverse=True)[:8]
        print("headline step parity, worst tensors:", [(k, f"{v:.2e}") for v, k in worst])
    tr.release()
    return errs


def attention_direct_case(M, device, C, H, N, T, seed=0, p_edge=0.3):
    """new_multihead_attention at the training step's sizes, DIRECTLY against O.attention (modules.py:236-311) on seeded random
    inputs: output, attention probabilities, input and parameter gradients.  Padded tokens (zero rows), an edge-less query and
    a query whose only edge is key-masked are included."""
    case = f"direct_attn_c{C}_n{N}_t{T}_{seed}"
    P0 = GS.make_params(case, GS.attention_shapes(C))
    x = GS.randn(f"{case}/x", N, T, C)
    x[:, T - 2:] = 0
    x[0, 3] = 0
    graph = GS.bernoulli(f"{case}/g", p_edge, N, T, T).float()
    graph[:, 1] = 0                      # no edge at all
    graph[:, 2] = 0
    graph[:, 2, T - 1] = 1               # only edge is a key-masked (padded) token
    dy = GS.randn(f"{case}/dy", N, T, C)
    ref = _oracle_attention(P0, x, x, graph, H, dy, None, True)
    emu = _oracle_attention(P0, x, x, graph, H, dy, BF, True)
    # probabilities from a return_att module (forward only); output and gradients from the module as the training step builds it
    # (return_att=False: forward statistics kept for the one-pass shared-tile backward kernel)
    m_att = M.new_multihead_attention(C, H, return_att=True)
    set_params(m_att, P0)
    m_att = m_att.to(device)
    with torch.no_grad():
        y_att, att = m_att(x.to(device), x.to(device), x.to(device), graph.to(device))
    m = M.new_multihead_attention(C, H)
    set_params(m, P0)
    m = m.to(device)
    xx = x.clone().to(device).requires_grad_(True)
    from savqa_b200 import ops as _ops
    gd = _ops.attach_graph_bits(graph.to(device))  # a 0/1 graph as AttModel_x3's mask builder hands it over: with its bit-packed form
    y = m(xx, xx, xx, gd)
    errs = {"y": check(f"{case}: output", y, ref["y"], emu["y"]), "y_att": check(f"{case}: output (return_att)", y_att, ref["y"], emu["y"]),
            "att": check(f"{case}: attention probabilities", att, ref["att"], emu["att"])}
    a4 = att.detach().cpu().view(H, N, T, T)
    assert float(a4[:, :, 1].abs().sum()) == 0.0 and float(a4[:, :, 2].abs().sum()) == 0.0
    (y * dy.to(device)).sum().backward()
    zero_q = (x.abs().sum(-1) == 0)
    errs["dx"] = check(f"{case}: d input", xx.grad, ref["dq"], emu["dq"], mask=zero_q)
    ours = grads_of(m, P0.keys())
    for k_ in P0:
        errs[k_] = check(f"{case}: grad {k_}", ours[k_], ref[k_], emu[k_])
    return errs


def deferred_adam_case(ops, device, steps=12, rows=40, width=24, seed=0):
    """savqa_adam_rows (deferred row-wise Adam) against DENSE torch.optim.Adam (main_itp_ddp_tar_super_node.py:206, 366) on a
    small table whose rows come and go between steps: after flush the tables agree, and every row that a step reads is current
    BEFORE the step reads it (what makes the forward pass see the dense optimizer's parameters)."""
    g = torch.Generator().manual_seed(seed)
    lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-8
    p0 = torch.randn(rows, width, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=lr, betas=(b1, b2), eps=eps)
    p = p0.clone().to(device)
    m, v = (torch.zeros(rows, width, device=device) for _ in range(2))
    grad = torch.zeros(rows, width, dtype=torch.int64, device=device)  # the trainer's Q15.48 accumulator (order-independent sums)
    stamp = torch.zeros(rows, dtype=torch.int32, device=device)
    dyn = torch.zeros(3, device=device)
    worst_read = 0.0
    for step in range(1, steps + 1):
        n = int(torch.randint(1, 9, (1,), generator=g))
        idx = torch.randint(0, rows, (n,), generator=g)
        idx = torch.cat([idx, idx[:2]])  # duplicates inside one step
        ops.adam_advance(dyn, lr, b1, b2)
        ops.adam_rows(p, None, m, v, stamp, idx.to(device), lr, b1, b2, eps, step, dyn=dyn, apply=False)  # catch-up before the read
        worst_read = max(worst_read, float((p[idx.to(device)].cpu() - ref.detach()[idx]).abs().max()))
        rows_g = torch.randn(idx.numel(), width, generator=g)
        dense = torch.zeros(rows, width)
        dense.index_add_(0, idx, rows_g)
        ref.grad = dense.clone()
        opt.step()
        ops.scatter_add_rows(grad, idx.to(device), rows_g.to(device))
        ops.adam_rows(p, grad, m, v, stamp, idx.to(device), lr, b1, b2, eps, step, dyn=dyn, apply=True)
        assert float(grad.abs().sum()) == 0.0  # consumed gradient rows are zeroed again
    ops.adam_rows(p, None, m, v, stamp, None, lr, b1, b2, eps, steps + 1, dyn=None, apply=False)  # flush: every row through `steps`
    err = float((p.cpu() - ref.detach()).abs().max())
    assert worst_read < 5e-6 and err < 5e-6, (worst_read, err)
    return err


def mil_nce_module_case(A, golden_dir, device, case):
    """Our MIL_NCE module (functional.MilNceFn: gather + tensor-core GEMMs + savqa_mil_nce_fwd/bwd) against the golden vectors of
    the live reference's MIL_NCE (AttModel_x3.py:285-443, only_obj) and the oracle: new_macro_ipt, mil_nce_obj, and the gradients
    of every parameter the forward uses for L = sum(out * w) + 3 * mil_nce_obj."""
    import types
    c = GS.MIL_CASES[case]
    g = load(golden_dir, case)
    P0 = GS.make_params(case, GS.mil_nce_shapes(c["h"]))
    b = GS.mil_nce_case(case, c["B"], c["V"], c["M"], c["topN"])
    w = GS.randn(f"{case}/dout", c["B"], c["M"], GS.F_REGION)

    def run(od):
        P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
        out, obj = O.mil_nce(P, b["vis_fea"], b["macro_ipt"], b["macro_obj_loc"], b["pos"], b["neg"], b["mask"], operand_dtype=od)
        ((out * w).sum() + 3.0 * obj).backward()
        return dict(out=out.detach(), obj=obj.detach(), **{k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()})

    ref, emu = run(None), run(BF)
    assert O.rel_err(ref["out"], t(g["out"])) < 2e-6 and abs(float(ref["obj"]) - float(g["obj"])) < 2e-6 * max(1.0, abs(float(g["obj"])))
    saved = A.VOCAB_ROWS
    A.VOCAB_ROWS = GS.SMALL_VOCAB
    try:
        m = A.MIL_NCE(types.SimpleNamespace(vectors=torch.zeros(4, GS.E_GLOVE)), c["h"], 0.0, 5, True)
    finally:
        A.VOCAB_ROWS = saved
    missing = m.load_state_dict({k: v.clone() for k, v in P0.items()}, strict=False)
    assert set(missing.missing_keys) == {"R", "rel_mlp.0.weight", "rel_mlp.0.bias", "rel_mlp.2.weight", "rel_mlp.2.bias", "bilinear.weight"}
    m = m.to(device)
    dv = lambda x: x.to(device)  # noqa: E731
    e = torch.empty((c["B"], 0), device=device)
    out, obj, rel = m(dv(b["vis_fea"]), dv(b["macro_ipt"]), dv(b["macro_obj_loc"]), dv(b["pos"]), dv(b["neg"]), dv(b["mask"]), e, e, e, e)
    assert out.shape == (c["B"], c["M"], GS.F_REGION) and out.dtype == torch.float32 and rel == 0
    errs = {"out": check(f"{case}: new_macro_ipt", out, t(g["out"]), emu["out"])}
    errs["obj"] = abs(float(obj) - float(ref["obj"])) / max(1.0, abs(float(ref["obj"])))
    assert errs["obj"] < 3e-3 + 3 * abs(float(emu["obj"]) - float(ref["obj"])) / max(1.0, abs(float(ref["obj"]))), errs
    ((out * w.to(device)).sum() + 3.0 * obj).backward()
    named = dict(m.named_parameters())
    for k in P0:
        got = named[k].grad
        if float(ref[k].abs().sum()) == 0.0:  # marco_mlp: `.detach()` at AttModel_x3.py:354
            assert got is None or float(got.abs().sum()) == 0.0, k
            continue
        errs[k] = check(f"{case}: grad {k}", got, ref[k], emu[k], floor=2e-2, factor=6.0)
    return errs


def compact_equals_dense_case(device, batch_size=6, cfg=None, vocab_rows=3000, train_mode=False):
    """AttModel.forward_compact on collate.compact_batch(batch) returns what AttModel.forward returns on the dense collate_fn batch
    whose features were rounded to bf16, BIT FOR BIT (the masks the kernels see are identical: savqa_build_masks_compact vs
    savqa_build_masks)."""
    from savqa_b200 import collate, synthetic
    cfg = cfg or dict(synthetic.GQA_SHAPED, V=12, Q=8, M=20, ncls=64)
    model = synthetic.build_model(cfg, vocab_rows=vocab_rows).to(device)
    model.train(train_mode)
    b = synthetic.make_batch(cfg, batch_size, seed=9, vocab_rows=vocab_rows)
    c = collate.compact_batch(b)
    dense = {k: v.to(device) for k, v in collate.expand_batch(c).items()}
    comp = {k: v.to(device) for k, v in c.items()}
    e = torch.empty((batch_size, 0), device=device)
    with torch.no_grad():
        ref = model(dense["vis_fea"], dense["vis_fea_mask"], dense["q_ipt"], dense["q_ipt_mask"], dense["q_ipt_graph"], dense["macro_node_ipt"],
                    dense["macro_node_mask"], dense["macro_graph_ipt"], dense["macro_obj_loc_ipt"], dense["micro_positive_obj_ipt"],
                    dense["micro_negative_obj_ipt"], dense["micro_obj_mask"], e, e, e, e, decMask=True, mcb=False)
        got = model.forward_compact(comp, decMask=True)
    for a_, b_ in zip(got[:4], ref[:4]):
        assert torch.equal(a_, b_)
    return got
