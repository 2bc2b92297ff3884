"""Host-side logic on CPU: autograd wiring of savqa_b200.functional and the module composition, with the C-ABI
kernels replaced by the torch stand-ins of tests/fake_ops.py (which restate the kernels' contracts, including their
explicit backward formulas).  Compared against the oracle at the bf16-operand tolerance.  Also: the drop-in
contract (state_dict keys, signatures) and the "fail loudly without the extension / a GPU" rule."""
import inspect
import json
import os

import numpy as np
import pytest
import torch

import fake_ops
from oracle import golden_spec as GS
from oracle import savqa_oracle as O


@pytest.fixture()
def fake(monkeypatch):
    fake_ops.install(monkeypatch)
    import savqa_b200.functional as Fn
    return Fn


import parity_cases as PC  # noqa: E402
from parity_util import check, grads_of, load, set_params, t  # noqa: E402,F401


@pytest.mark.parametrize("tag,C,H,N,T", [("c64", 64, 4, 3, 10), ("c512", 512, 8, 2, 24)])
def test_attention_self_wiring(fake, golden_dir, tag, C, H, N, T):
    from savqa_b200 import modules as M
    PC.attention_case(M, golden_dir, "cpu", f"attn_self_{tag}", C, H, N, T, T, True)


@pytest.mark.parametrize("tq", [1, 3])
def test_attention_cross_wiring(fake, golden_dir, tq):
    from savqa_b200 import modules as M
    PC.attention_case(M, golden_dir, "cpu", f"attn_cross{tq}_c64", 64, 4, 3, tq, 10, False)


def test_mha_causal_and_graphmask_wiring(fake, golden_dir):
    from savqa_b200 import modules as M
    PC.attention_case(M, golden_dir, "cpu", "mha_causal_c64", 64, 4, 3, 6, 6, True, kind="mha")
    PC.attention_case(M, golden_dir, "cpu", "mha_token_c64", 64, 4, 5, 1, 1, True, kind="mha")  # one token: the single-GEMM form
    PC.attention_case(M, golden_dir, "cpu", "attn_graphmask_c64", 64, 4, 3, 10, 10, True, kind="gm")
    m = M.new_multihead_attention_with_graph_mask(64, 4)
    q = torch.randn(2, 3, 64)
    with pytest.raises(AttributeError):
        m(q, q, q, None, None)  # the reference fails the same way (modules.py:375)


@pytest.mark.parametrize("tag,C,N,T", [("c64", 64, 3, 10), ("c512", 512, 2, 24)])
def test_feedforward_wiring(fake, golden_dir, tag, C, N, T):
    from savqa_b200 import modules as M
    PC.feedforward_case(M, golden_dir, "cpu", f"ffn_{tag}", C, N, T)


def test_layernorm_and_embedding_wiring(fake, golden_dir):
    from savqa_b200 import modules as M
    PC.layernorm_case(M, golden_dir, "cpu")
    PC.embedding_cases(M, golden_dir, "cpu")


@pytest.mark.parametrize("kind", ["vis", "syb"])
def test_branch_wiring(fake, golden_dir, kind):
    from savqa_b200 import AttModel_x3 as A
    PC.branch_case(A, golden_dir, "cpu", kind)


def test_heads_and_loss_wiring(fake, golden_dir):
    from savqa_b200 import AttModel_x3 as A
    S = GS.SMALL
    g = load(golden_dir, "full_c64")
    Ph = GS.make_params("full_c64", GS.head_shapes(S["C"], S["ncls"]))
    heads = torch.nn.Module()
    model = A.AttModel.__new__(A.AttModel)
    torch.nn.Module.__init__(model)
    model.cls = A._head(2 * S["C"], S["C"], S["ncls"], 0.0)
    model.cls_vis = A._head(S["C"], S["C"], S["ncls"], 0.0)
    model.cls_syb = A._head(S["C"], S["C"], S["ncls"], 0.0)
    model._pk = {k: (A.WeightPack(), A.WeightPack()) for k in ("cls", "cls_vis", "cls_syb")}
    model.load_state_dict(Ph, strict=True)
    # decoder outputs from the oracle, logits through our heads
    Pv = GS.make_params("branch_vis_c64", GS.branch_shapes("vis", S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]))
    Ps = GS.make_params("branch_syb_c64", GS.branch_shapes("syb", S["C"], S["maxlen"], S["maxlen_q"], S["maxlen_v"], S["blocks"], S["ncls"]))
    bv = GS.branch_case("branch_vis_c64", "vis", S["B"], S["V"], S["Q"])
    bs = GS.branch_case("branch_syb_c64", "syb", S["B"], S["M"], S["Q"])
    fv = O.branch_forward(Pv, "vis", bv["first"], bv["first_mask"], None, bv["q_ipt"], bv["q_graph"], bv["q_mask"], True, S["blocks"], S["heads"])
    fs = O.branch_forward(Ps, "syb", t(g["syb_ipt"]), bs["first_mask"], bs["first_graph"], bv["q_ipt"], bv["q_graph"], bv["q_mask"], True,
                          S["blocks"], S["heads"])
    lc, lv, ls = model.answer_logits(fv, fs)
    assert O.rel_err(lc, t(g["logits_concat"])) < 1e-2
    assert O.rel_err(lv, t(g["logits_vis"])) < 1e-2
    answer = GS.randint("full_c64/answer", 0, S["ncls"], S["B"])
    loss = A.answer_loss(t(g["logits_concat"]).clone().requires_grad_(True), t(g["logits_vis"]), t(g["logits_syb"]), answer)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    # gradient of the fused loss vs autograd through the oracle
    L = [t(g[k]).clone().requires_grad_(True) for k in ("logits_concat", "logits_vis", "logits_syb")]
    O.answer_loss(*L, answer).backward()
    L2 = [t(g[k]).clone().requires_grad_(True) for k in ("logits_concat", "logits_vis", "logits_syb")]
    A.answer_loss(*L2, answer).backward()
    for a, b_ in zip(L, L2):
        assert O.rel_err(b_.grad, a.grad) < 1e-5


def test_signatures_match_reference():
    """Constructor / forward argument names of the drop-in classes."""
    from savqa_b200 import modules as M, AttModel_x3 as A

    def args(f):
        return list(inspect.signature(f).parameters)[1:]

    assert args(M.new_multihead_attention.__init__) == ["num_units", "num_heads", "dropout_rate", "causality", "return_att"]
    assert args(M.new_multihead_attention.forward) == ["queries", "keys", "values", "graph"]
    assert args(M.multihead_attention.__init__) == ["num_units", "num_heads", "dropout_rate", "causality"]
    assert args(M.multihead_attention.forward) == ["queries", "keys", "values"]
    assert args(M.new_multihead_attention_with_graph_mask.forward) == ["queries", "keys", "values", "key_mask_ipt", "graph"]
    assert args(M.feedforward.__init__) == ["in_channels", "num_units"]
    assert args(M.layer_normalization.__init__) == ["features", "epsilon"]
    assert args(M.embedding.__init__) == ["vocab_size", "num_units", "zeros_pad", "scale"]
    assert args(A.AttModel_vis_grid.__init__) == ["glove", "hidden_size", "maxlen", "maxlen_q", "num_blocks", "num_heads", "dropout_rate",
                                                  "maxlen_v", "num_classes"]
    assert args(A.AttModel_vis_grid.forward) == ["vis_fea", "vis_mask", "q_fea", "q_graph", "q_mask", "decMask"]
    assert args(A.AttModel_syb.__init__) == ["glove", "hidden_size", "maxlen", "maxlen_q", "num_blocks", "num_heads", "dropout_rate", "num_classes"]
    assert args(A.AttModel_syb.forward) == ["syb_ipt", "syb_mask", "syb_graph", "q_fea", "q_graph", "q_mask", "decMask"]
    assert args(A.AttModel.__init__) == ["glove", "hidden_size", "hidden_size_mil", "num_classes", "maxlen_q", "maxlen", "maxlen_v", "num_blocks",
                                         "num_heads", "dropout_rate", "dropout_rate_mcb", "num_relations", "only_obj"]
    assert args(A.AttModel.forward)[:5] == ["vis_fea", "vis_mask", "q_ipt", "q_mask", "q_graph"]
    assert args(A.AttModel.forward)[-2:] == ["decMask", "mcb"]


def test_full_state_dict_contract(golden_dir):
    """Every key/shape of the reference AttModel state_dict (dumped from the live reference) exists in ours."""
    import types
    from savqa_b200 import AttModel_x3 as A
    S = GS.SMALL
    keys = {k: tuple(s) for k, s in json.load(open(os.path.join(golden_dir, "state_dict_keys_c64.json")))}
    glove = types.SimpleNamespace(vectors=torch.zeros(4, 300))
    m = A.AttModel(glove, S["C"], 16, S["ncls"], S["maxlen_q"], S["maxlen"], S["maxlen_v"], S["blocks"], S["heads"], 0.0, 0.1, 5, True)
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ours == keys, (sorted(set(ours) ^ set(keys))[:8], [k for k in keys if k in ours and ours[k] != keys[k]][:8])


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: on a box without CUDA every op of the product path raises (never computes on the host)."""
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from savqa_b200 import modules as M
    m = M.feedforward(64, [256, 64])
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(torch.randn(2, 3, 64))
    a = M.new_multihead_attention(64, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        x = torch.randn(2, 3, 64)
        a(x, x, x, torch.ones(2, 3, 3))


@pytest.mark.parametrize("case", ["mil_nce_h16_top2", "mil_nce_h64_top1", "mil_nce_h128_top5"])
def test_mil_nce_wiring_vs_golden(monkeypatch, golden_dir, case):
    """functional.MilNceFn (forward and hand-written backward) with the kernels replaced by their CPU restatements, against the
    live reference's MIL_NCE golden and the oracle."""
    fake_ops.install(monkeypatch)
    from savqa_b200 import AttModel_x3 as A
    errs = PC.mil_nce_module_case(A, golden_dir, "cpu", case)
    assert errs["out"] < 1e-2


def test_compact_hand_off_equals_dense(monkeypatch):
    """collate.compact_batch -> AttModel.forward_compact == AttModel.forward on the dense batch, bit for bit; the compact batch is
    ~8x smaller than the dense one."""
    fake_ops.install(monkeypatch)
    from savqa_b200 import collate, synthetic
    PC.compact_equals_dense_case("cpu")
    b = synthetic.make_batch(synthetic.GQA_SHAPED, 4, seed=1, vocab_rows=2000)
    c = collate.compact_batch(b)
    nbytes = lambda d, keys: sum(d[k].numel() * d[k].element_size() for k in keys)  # noqa: E731
    from savqa_b200 import train
    assert nbytes(c, collate.COMPACT_KEYS) * 2.5 < nbytes(b, train.FULL_KEYS)
    bad = dict(b)
    bad["vis_fea_mask"] = b["vis_fea_mask"].clone()
    bad["vis_fea_mask"][0, 0, 1] = 0
    with pytest.raises(ValueError):
        collate.compact_batch(bad)


def test_evaluate_loop_matches_reference_eval(monkeypatch):
    """infer.evaluate == the reference's eval() for --model_v 3 (main_itp_ddp_tar_super_node.py:60-142): batch-size-weighted mean of the
    label-smoothed 3-head loss (+ the MIL-NCE term), accuracy over the samples whose answer id is not 0 -- here against the oracle's
    full_step on the same batches, from collate_fn's dense batches and from the compact hand-off."""
    fake_ops.install(monkeypatch)
    from savqa_b200 import collate, infer, synthetic
    cfg = synthetic.TINY
    model = synthetic.build_model(cfg, vocab_rows=1200)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    dense = [collate.expand_batch(collate.compact_batch(synthetic.make_batch(cfg, n, seed=40 + n, vocab_rows=1200))) for n in (4, 6)]
    dense[0]["answer"][0] = 0  # never counted as correct (main...:125)
    tot = cnt = correct = 0.0
    for b in dense:
        with torch.no_grad():
            loss, logits, obj, _ = O.full_step(sd, b, cfg["blocks"], cfg["heads"], with_milnce_loss=True)
        lsm = sum(torch.log_softmax(l, -1) for l in logits) / 3
        pred = lsm.argmax(1)
        n = b["answer"].shape[0]
        tot += float(loss) * n
        cnt += n
        correct += float(((pred == b["answer"]) & (b["answer"] != 0)).sum())
    for batches in (dense, [collate.compact_batch(b) for b in dense]):
        loss, ok, n = infer.evaluate(model, batches, dec_mask=True, with_milnce_loss=True)
        assert n == cnt and abs(loss - tot / cnt) < 5e-3 * abs(tot / cnt), (loss, tot / cnt)
        assert abs(ok - correct) <= 1  # bf16 operands may flip one near-tie argmax


def test_fixed_point_row_accumulator_is_order_independent():
    """The word tables' gradient accumulator (savqa_scatter_add_rows_q48, here its CPU stand-in with the same rounding): the sum of
    a list of rows with many duplicates does not depend on the order of the additions -- what keeps data-parallel replicas
    bit-identical (float accumulation does depend on it) -- and equals the float64 sum to 2^-48 per addend."""
    import fake_ops as F
    g = torch.Generator().manual_seed(11)
    n, rows, width = 3000, 13, 24
    idx = torch.randint(0, rows, (n,), generator=g)
    idx[::17] = 5  # one very hot row
    grad_rows = torch.randn(n, width, generator=g) * 1e-3
    perm = torch.randperm(n, generator=g)
    a = torch.zeros(rows, width, dtype=torch.int64)
    b = torch.zeros(rows, width, dtype=torch.int64)
    F.scatter_add_rows(a, idx, grad_rows, scale=0.125, skip_row=7)
    F.scatter_add_rows(b, idx[perm], grad_rows[perm], scale=0.125, skip_row=7)
    assert torch.equal(a, b)
    assert int(a[7].abs().sum()) == 0  # the padding row takes no gradient
    ref = torch.zeros(rows, width, dtype=torch.float64)
    keep = idx != 7
    ref.index_add_(0, idx[keep], (grad_rows[keep] * 0.125).double())
    assert float((a.double() / 2.0 ** 48 - ref).abs().max()) <= n * 2.0 ** -48
    # float accumulation of the same two orders differs in the last bits (the reason for the fixed-point form)
    fa, fb = torch.zeros(rows, width), torch.zeros(rows, width)
    for i in range(n):
        fa[idx[i]] += grad_rows[i] * 0.125
    for i in perm.tolist():
        fb[idx[i]] += grad_rows[i] * 0.125
    assert not torch.equal(fa, fb)
