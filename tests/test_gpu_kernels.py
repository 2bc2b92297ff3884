"""GPU parity tests, kernel level: every C-ABI entry point against the oracle / a plain fp32 restatement on the same
seeded inputs.  Integer / index / mask work is compared bit-exactly; floating point with the tolerance written
next to each assert.  Run on the B200 box with `-m gpu`."""
import math

import pytest
import torch

from oracle import golden_spec as GS
from oracle import savqa_oracle as O

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    from savqa_b200 import _lib, ops as o
    _lib.require_device()
    return o


def dev(x):
    return x.cuda() if x is not None else None


def rel(a, b):
    return O.rel_err(a.float().cpu(), b.float().cpu())


# ------------------------------------------------------------------------------------------------ masks
@pytest.mark.parametrize("B,V,Q,with_first_graph,as_float", [(3, 5, 4, False, False), (4, 36, 20, True, False), (2, 100, 20, True, True),
                                                            (2, 1, 1, True, False), (128, 36, 20, False, False)])
def test_build_masks_bit_exact(ops, B, V, Q, with_first_graph, as_float):
    b = GS.branch_case(f"masks/{B}/{V}/{Q}", "syb" if with_first_graph else "vis", B, V, Q)
    cast = (lambda x: x.float()) if as_float else (lambda x: x)
    fm, qm, qg = cast(b["first_mask"]), cast(b["q_mask"]), cast(b["q_graph"])
    fg = cast(b["first_graph"]) if with_first_graph else None
    for dec_on in (True, False):
        ref = O.build_masks(fm, qm, qg, fg, dec_on)
        got = ops.build_masks(dev(fm), dev(qm), dev(qg), dev(fg), dec_on)
        for r, g_ in zip(ref, got):
            assert torch.equal(r, g_.cpu())


# ------------------------------------------------------------------------------------------------ gather / scatter
def test_gather_scatter(ops):
    table = GS.randn("gk/table", 1000, 300)
    idx = GS.randint("gk/idx", 0, 1000, 7, 20)
    idx[0, :3] = torch.tensor([0, 999, 999])
    o32, o16 = ops.gather_rows(dev(table), dev(idx), want_f32=True, want_bf16=True)
    assert torch.equal(o32.cpu(), table[idx.reshape(-1)])  # bit exact
    assert o16.shape == (140, 304)
    assert torch.equal(o16.cpu()[:, :300], table[idx.reshape(-1)].to(BF))
    assert float(o16[:, 300:].float().abs().sum()) == 0.0
    o32s, _ = ops.gather_rows(dev(table), dev(idx), scale=math.sqrt(300.0))
    assert torch.equal(o32s.cpu(), table[idx.reshape(-1)] * torch.tensor(math.sqrt(300.0), dtype=F32))
    # scatter-add with a padding row that must not receive gradient
    dout = GS.randn("gk/dout", 140, 300)
    dt = torch.zeros(1000, 300, device="cuda")
    ops.scatter_add_rows(dt, dev(idx), dev(dout), scale=1.0, skip_row=999)
    ref = torch.zeros(1000, 300)
    flat = idx.reshape(-1)
    keep = flat != 999
    ref.index_add_(0, flat[keep], dout[keep])
    assert rel(dt, ref) < 1e-6
    assert float(dt[999].abs().sum()) == 0.0
    # the fixed-point accumulator: heavily duplicated rows, two runs -> the same bits (what keeps data-parallel replicas identical),
    # and the value of the float64 sum to 2^-48 per addend
    hot = torch.randint(0, 7, (4000,), generator=torch.Generator().manual_seed(3))
    hot[::50] = 999
    big = GS.randn("gk/dout_hot", 4000, 300) * 1e-3
    q = [torch.zeros(1000, 300, dtype=torch.int64, device="cuda") for _ in range(2)]
    for t in q:
        ops.scatter_add_rows(t, dev(hot), dev(big), scale=0.5, skip_row=999)
    assert torch.equal(q[0], q[1])
    ref64 = torch.zeros(1000, 300, dtype=torch.float64)
    ref64.index_add_(0, hot[hot != 999], (big * 0.5).double()[hot != 999])
    assert float((q[0].cpu().double() / ops.Q48 - ref64).abs().max()) < 4000 * 2.0 ** -48
    assert int(q[0][999].abs().sum()) == 0
    # empty input is a no-op
    ops.gather_rows(dev(table), torch.zeros(0, dtype=torch.int64, device="cuda"))


# ------------------------------------------------------------------------------------------------ casts / masks / sums
def test_casts_rowmask_colsum(ops):
    x = GS.randn("ck/x", 77, 300)
    x[5] = 0
    xb = ops.cast_bf16(dev(x))
    assert xb.shape == (77, 304) and torch.equal(xb.cpu()[:, :300], x.to(BF)) and float(xb[:, 300:].float().abs().sum()) == 0
    w = GS.randn("ck/w", 130, 70)
    wt = torch.zeros(70, 136, dtype=BF, device="cuda")
    ops.cast_transpose_bf16(dev(w), wt)
    assert torch.equal(wt.cpu()[:, :130], w.t().to(BF))
    on, xb2 = ops.row_nonzero(dev(x))
    assert torch.equal(on.cpu(), (x.sum(-1) != 0).float()) and on[5] == 0
    assert torch.equal(xb2.cpu()[:, :300], x.to(BF))
    y = GS.randn("ck/y", 1000, 200).to(BF)
    out = torch.zeros(200, device="cuda")
    ops.colsum_bf16(dev(y), out)
    assert rel(out, y.float().sum(0)) < 1e-5
    act = GS.randn("ck/act", 77, 304).to(BF)
    gated = ops.relu_gate_bf16(dev(x), dev(act))
    assert torch.equal(gated.cpu()[:, :300], torch.where(act[:, :300].float() > 0, x, torch.zeros(())).to(BF))


# ------------------------------------------------------------------------------------------------ layer norm
@pytest.mark.parametrize("rows,C", [(35, 512), (1000, 512), (33, 64), (17, 1024), (9, 300)])
def test_layernorm_kernels(ops, rows, C):
    x = GS.randn(f"ln/{rows}/{C}/x", rows, C, scale=2.0)
    r = GS.randn(f"ln/{rows}/{C}/r", rows, C)
    x[0] = 0.75
    r[0] = 0.5  # constant row: sigma == 0
    gamma = GS.rand(f"ln/{C}/g", C, lo=0.8, hi=1.2)
    beta = GS.randn(f"ln/{C}/b", C, scale=0.1)
    dy = GS.randn(f"ln/{rows}/{C}/dy", rows, C)
    pre = (x + r).clone().requires_grad_(True)
    gg, bb = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = O.layer_norm(pre, gg, bb)
    (y_ref * dy).sum().backward()
    y, pre_k, yb, on = ops.layernorm_fwd(dev(x), dev(r), dev(gamma), dev(beta), 1e-8, True, True, True)
    assert rel(y, y_ref) < 2e-6  # fp32 arithmetic, different reduction order only
    assert torch.equal(pre_k.cpu(), x + r)
    assert torch.equal(yb.cpu(), y.cpu().to(BF))
    assert torch.equal(on.cpu(), (y.cpu().sum(-1) != 0).float())
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dsum = torch.zeros(C, device="cuda")
    dx, dxb = ops.layernorm_bwd(dev(dy), pre_k, dev(gamma), 1e-8, dg, db, want_bf16=True, dxsum=dsum)
    assert float((dsum - dx.sum(0)).abs().max()) <= 1e-5 * float(dx.abs().max()) * rows  # column sums of dx (bias gradient)
    assert rel(dx[1:], pre.grad[1:]) < 1e-5
    assert rel(dx[0], pre.grad[0]) < 1e-4  # sigma == 0 row: (g - mean g) / eps, as autograd
    assert rel(dg, gg.grad) < 1e-4 and rel(db, bb.grad) < 1e-5
    assert torch.equal(dxb.cpu(), dx.cpu().to(BF))


# ------------------------------------------------------------------------------------------------ GEMM (tcgen05)
def _gemm_ref(a, b, bias=None, res=None, rowtab=None, period=0, relu=False, gate=None, alpha=1.0):
    v = alpha * (a.float() @ b.float().t())
    if bias is not None:
        v = v + bias
    if res is not None:
        v = v + res
    if rowtab is not None:
        v = v + rowtab[:period].repeat((v.shape[0] + period - 1) // period, 1)[: v.shape[0]]
    if relu:
        v = torch.relu(v)
    if gate is not None:
        v = v * (gate.float() > 0)
    return v


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 512), (200, 136, 304), (56 * 8, 1536, 512), (7168, 2048, 512),
                                   (7168, 512, 2048), (130, 48, 512), (128, 1845, 512), (1, 512, 512), (4096, 6144, 512)])
def test_gemm_kmajor(ops, M, N, K):
    a = GS.randn(f"gemm/{M}/{N}/{K}/a", M, K).to(BF)
    b = (GS.randn(f"gemm/{M}/{N}/{K}/b", N, K) / math.sqrt(K)).to(BF)
    bias = GS.randn(f"gemm/{N}/bias", N)
    ref = _gemm_ref(a, b, bias=bias, relu=True)
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=BF) if N % 8 == 0 else None
    ops.gemm(dev(a), dev(b), M, N, K, bias=dev(bias), relu=True, out_f32=o32, out_bf16=o16)
    torch.cuda.synchronize()
    assert rel(o32, ref) < 2e-5  # identical bf16 operands, fp32 accumulation: summation order only
    if o16 is not None:
        assert rel(o16, ref) < 4e-3  # + one bf16 rounding of the output


def test_gemm_epilogues(ops):
    M, N, K, T = 300, 512, 2048, 60
    a = GS.randn("gemm/e/a", M, K).to(BF)
    b = (GS.randn("gemm/e/b", N, K) / math.sqrt(K)).to(BF)
    bias, res = GS.randn("gemm/e/bias", N), GS.randn("gemm/e/res", M, N)
    rowtab = GS.randn("gemm/e/rt", 64, N)
    gate = GS.randn("gemm/e/gate", M, N).to(BF)
    ref = _gemm_ref(a, b, bias=bias, res=res, rowtab=rowtab, period=T, gate=gate, alpha=0.5)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(dev(a), dev(b), M, N, K, bias=dev(bias), res=dev(res), rowtab=dev(rowtab), rowtab_period=T, gate=dev(gate), alpha=0.5, out_f32=out)
    assert rel(out, ref) < 2e-5
    # accumulate modes
    base = GS.randn("gemm/e/base", M, N)
    acc = dev(base.clone())
    ops.gemm(dev(a), dev(b), M, N, K, out_f32=acc, accumulate=1)
    assert rel(acc, base + _gemm_ref(a, b)) < 2e-5
    acc = dev(base.clone())
    ops.gemm(dev(a), dev(b), M, N, K, out_f32=acc, accumulate=2, split_k=4)
    assert rel(acc, base + _gemm_ref(a, b)) < 2e-5
    # strided views (fused QKV layout): B operand and output are column slices
    big = torch.zeros(M, 3 * N, device="cuda", dtype=BF)
    ops.gemm(dev(a), dev(b), M, N, K, out_bf16=big[:, N:2 * N])
    assert rel(big[:, N:2 * N], _gemm_ref(a, b)) < 4e-3 and float(big[:, :N].float().abs().sum()) == 0


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 128, 128), (7168, 1536, 512), (16384, 2048, 512), (1000, 512, 2048), (2560, 2048, 304),
                                   (300, 192, 512), (7168, 1024, 512)])
def test_gemm_pair_kernel_forward_epilogues(ops, M, N, K):
    """The CTA-pair kernel (cta_group::2, TMA-store epilogue; taken for M > 128, N % 64 == 0, one output): bias + ReLU -> bf16,
    bias + residual -> fp32, bias + positional row table -> fp32, with ragged M (rows past M are clipped by the TMA store)."""
    a = GS.randn(f"gemm2/{M}/{K}/a", M, K).to(BF)
    b = (GS.randn(f"gemm2/{N}/{K}/b", N, K) / math.sqrt(K)).to(BF)
    bias = GS.randn(f"gemm2/{N}/bias", N)
    res = GS.randn(f"gemm2/{M}/{N}/res", M, N)
    rowtab = GS.randn(f"gemm2/{N}/rt", 64, N)
    o16 = torch.full((M + 3, N), 7.0, device="cuda", dtype=BF)
    ops.gemm(dev(a), dev(b), M, N, K, bias=dev(bias), relu=True, out_bf16=o16)
    torch.cuda.synchronize()
    assert rel(o16[:M], _gemm_ref(a, b, bias=bias, relu=True)) < 4e-3
    assert float((o16[M:].float() - 7.0).abs().max()) == 0.0  # nothing written past row M
    o32 = torch.full((M + 3, N), 7.0, device="cuda")
    ops.gemm(dev(a), dev(b), M, N, K, bias=dev(bias), res=dev(res), out_f32=o32)
    torch.cuda.synchronize()
    assert rel(o32[:M], _gemm_ref(a, b, bias=bias, res=res)) < 2e-5
    assert float((o32[M:] - 7.0).abs().max()) == 0.0
    ops.gemm(dev(a), dev(b), M, N, K, bias=dev(bias), rowtab=dev(rowtab), rowtab_period=56, alpha=0.5, out_f32=o32)
    torch.cuda.synchronize()
    assert rel(o32[:M], _gemm_ref(a, b, bias=bias, rowtab=rowtab, period=56, alpha=0.5)) < 2e-5
    # column-slice output of a wider buffer (fused QKV layout)
    big = torch.zeros(M, 3 * N, device="cuda", dtype=BF)
    ops.gemm(dev(a), dev(b), M, N, K, out_bf16=big[:, N:2 * N])
    torch.cuda.synchronize()
    assert rel(big[:, N:2 * N], _gemm_ref(a, b)) < 4e-3 and float(big[:, :N].float().abs().sum()) == 0 and float(big[:, 2 * N:].float().abs().sum()) == 0


@pytest.mark.parametrize("M0,M1,N,K", [(7168, 16384, 1536, 512), (7168, 16384, 512, 2048), (1000, 300, 256, 128), (2560, 2560, 2048, 304)])
def test_gemm_grouped_two_problems(ops, M0, M1, N, K):
    """One launch for the same layer of the two branch models (different row counts and weights): every output equals the
    single-problem launch bit for bit (same K order per tile); split-K weight gradients agree to fp32 summation order."""
    P = []
    for i, M in enumerate((M0, M1)):
        a = GS.randn(f"gg/{i}/{M}/{K}/a", M, K).to(BF)
        b = (GS.randn(f"gg/{i}/{N}/{K}/b", N, K) / math.sqrt(K)).to(BF)
        bias = GS.randn(f"gg/{i}/{N}/bias", N)
        res = GS.randn(f"gg/{i}/{M}/{N}/res", M, N)
        gate = GS.randn(f"gg/{i}/{M}/{N}/gate", M, N).to(BF)
        P.append(dict(a=dev(a), b=dev(b), bias=dev(bias), res=dev(res), gate=dev(gate), M=M))
    # forward kind: bias + ReLU -> bf16
    outs = [torch.empty(p["M"], N, device="cuda", dtype=BF) for p in P]
    ops.gemm_grouped([dict(a=p["a"], b=p["b"], M=p["M"], K=K, bias=p["bias"], relu=True, out_bf16=o) for p, o in zip(P, outs)], N)
    for p, o in zip(P, outs):
        single = torch.empty_like(o)
        ops.gemm(p["a"], p["b"], p["M"], N, K, bias=p["bias"], relu=True, out_bf16=single)
        assert torch.equal(o, single)
        assert rel(o, _gemm_ref(p["a"].cpu(), p["b"].cpu(), bias=p["bias"].cpu(), relu=True)) < 4e-3
    # residual kind -> fp32
    outs = [torch.empty(p["M"], N, device="cuda") for p in P]
    ops.gemm_grouped([dict(a=p["a"], b=p["b"], M=p["M"], K=K, bias=p["bias"], res=p["res"], out_f32=o) for p, o in zip(P, outs)], N)
    for p, o in zip(P, outs):
        single = torch.empty_like(o)
        ops.gemm(p["a"], p["b"], p["M"], N, K, bias=p["bias"], res=p["res"], out_f32=single)
        assert torch.equal(o, single)
    # dgrad kind: dX = dY W (W read MN-major), ReLU gate + column sums
    if K % 64 == 0:
        outs = [torch.empty(p["M"], K, device="cuda", dtype=BF) for p in P]
        cs = [torch.zeros(K, device="cuda") for _ in P]
        dys = [dev(GS.randn(f"gg/{i}/dy", p["M"], N).to(BF)) for i, p in enumerate(P)]
        gates = [dev(GS.randn(f"gg/{i}/g2", p["M"], K).to(BF)) for i, p in enumerate(P)]
        ops.gemm_grouped([dict(a=dy, b=p["b"], M=p["M"], K=N, gate=g_, out_bf16=o, colsum=c) for p, dy, g_, o, c in zip(P, dys, gates, outs, cs)],
                         K, b_mn=True)
        for p, dy, g_, o, c in zip(P, dys, gates, outs, cs):
            single, c1 = torch.empty_like(o), torch.zeros(K, device="cuda")
            ops.gemm(dy, p["b"], p["M"], K, N, b_mn=True, gate=g_, out_bf16=single, colsum=c1)
            assert torch.equal(o, single) and rel(c, c1) < 1e-5
    # wgrad kind
    dws = [torch.zeros(N, K, device="cuda") for _ in P]
    dys = [dev(GS.randn(f"gg/{i}/dyw", p["M"], N).to(BF)) for i, p in enumerate(P)]
    ops.wgrad_grouped([(dy, p["a"], N, K, dw) for p, dy, dw in zip(P, dys, dws)])
    for p, dy, dw in zip(P, dys, dws):
        assert rel(dw, dy.float().t() @ p["a"].float()) < 2e-5


@pytest.mark.parametrize("M,Nout,Kin", [(300, 512, 2048), (7168, 2048, 512), (128, 1845, 512), (140, 2048, 300), (130, 1536, 512),
                                        (16384, 512, 2048), (16384, 1536, 512), (1000, 2048, 512)])
def test_gemm_dgrad_mn_major_b_and_colsum(ops, M, Nout, Kin):
    """dX[M,Kin] = dY[M,Nout] W[Nout,Kin] with W read MN-major from its forward staging [Nout, pad8(Kin)] (no transposed copy),
    ReLU gate and the bias-gradient column sums fused in the epilogue (functional.dgrad)."""
    ldw = ops.pad8(Kin)
    w = torch.zeros(Nout, ldw, dtype=BF)
    w[:, :Kin] = (GS.randn(f"dg/{Nout}/{Kin}/w", Nout, Kin) / math.sqrt(Nout)).to(BF)
    dy = torch.zeros(M, ops.pad8(Nout), dtype=BF)
    dy[:, :Nout] = GS.randn(f"dg/{M}/{Nout}/dy", M, Nout).to(BF)
    gate = GS.randn(f"dg/{M}/{Kin}/gate", M, ldw).to(BF)
    ref = (dy[:, :Nout].float() @ w[:, :Kin].float()) * (gate[:, :Kin].float() > 0)
    out = torch.empty(M, Kin, device="cuda")
    outb = torch.empty(M, ldw, device="cuda", dtype=BF)
    cs = torch.zeros(Kin, device="cuda")
    ops.gemm(dev(dy), dev(w), M, Kin, Nout, b_mn=True, gate=dev(gate), out_f32=out, out_bf16=outb, colsum=cs)
    torch.cuda.synchronize()
    assert rel(out, ref) < 2e-5
    assert rel(outb[:, :Kin], ref) < 4e-3
    assert rel(cs, ref.sum(0)) < 1e-4  # fp32 atomics over the row tiles: order only
    # single-output forms (the ones functional.dgrad issues; these run on the CTA-pair kernel when M > 128 and Kin % 64 == 0)
    cs2 = torch.zeros(Kin, device="cuda")
    outb2 = torch.zeros(M, ldw, device="cuda", dtype=BF)
    ops.gemm(dev(dy), dev(w), M, Kin, Nout, b_mn=True, gate=dev(gate), out_bf16=outb2, colsum=cs2)
    res = GS.randn(f"dg/{M}/{Kin}/res", M, Kin)
    out2 = torch.empty(M, Kin, device="cuda")
    ops.gemm(dev(dy), dev(w), M, Kin, Nout, b_mn=True, res=dev(res), out_f32=out2)
    torch.cuda.synchronize()
    assert rel(outb2[:, :Kin], ref) < 4e-3 and rel(cs2, ref.sum(0)) < 1e-4
    assert rel(out2, dy[:, :Nout].float() @ w[:, :Kin].float() + res) < 2e-5


@pytest.mark.parametrize("Mtok,Nout,Kin", [(256, 128, 128), (448, 1536, 512), (7168, 2048, 512), (7168, 512, 2048), (140, 2048, 300),
                                           (128, 1845, 512), (16384, 1536, 512), (16384, 512, 2048), (1000, 1024, 512)])
def test_gemm_wgrad_mn_major(ops, Mtok, Nout, Kin):
    """dW[Nout,Kin] = dY[Mtok,Nout]^T X[Mtok,Kin]: both operands MN-major, split-K with fp32 atomics."""
    ldy, ldx = ops.pad8(Nout), ops.pad8(Kin)
    dy = torch.zeros(Mtok, ldy, dtype=BF)
    x = torch.zeros(Mtok, ldx, dtype=BF)
    dy[:, :Nout] = GS.randn(f"wg/{Mtok}/{Nout}/dy", Mtok, Nout).to(BF)
    x[:, :Kin] = GS.randn(f"wg/{Mtok}/{Kin}/x", Mtok, Kin).to(BF)
    ref = dy[:, :Nout].float().t() @ x[:, :Kin].float()
    out = torch.zeros(Nout, Kin, device="cuda")
    ops.wgrad(dev(dy), dev(x), Nout, Kin, out)
    assert rel(out, ref) < 2e-5


# ------------------------------------------------------------------------------------------------ attention core
def _attn_inputs(tag, N, H, Tq, Tk, d, self_att):
    C = H * d
    q = torch.relu(GS.randn(f"{tag}/q", N * Tq, C)).to(BF)
    k = q if self_att else torch.relu(GS.randn(f"{tag}/k", N * Tk, C)).to(BF)
    v = torch.relu(GS.randn(f"{tag}/v", N * Tk, C)).to(BF)
    graph = GS.bernoulli(f"{tag}/g", 0.3, N, Tq, Tk).float()
    key_on = torch.ones(N, Tk)
    query_on = torch.ones(N, Tq)
    if Tk > 3:
        key_on[0, Tk - 1] = 0
        key_on[N - 1, 1] = 0
    if Tq > 3:
        graph[0, 1] = 0
        graph[0, 2] = 0
        graph[0, 2, Tk - 1] = 1
        query_on[N - 1, 0] = 0
    return q, k, v, graph, key_on, query_on


def _attn_ref(q, k, v, graph, key_on, query_on, N, H, Tq, Tk, d, causal, renorm):
    import fake_ops
    return fake_ops.graph_attention_fwd(q, k, v, graph, key_on, query_on, N, H, Tq, Tk, d, causal, renorm, True, 1)


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("N,H,Tq,Tk,d,renorm,causal", [
    (3, 8, 56, 56, 64, 1, False), (2, 8, 120, 120, 64, 1, False), (2, 8, 299, 299, 64, 1, False), (2, 4, 130, 130, 128, 1, False),
    (4, 8, 1, 56, 64, 1, False), (2, 8, 36, 36, 64, 0, True), (2, 8, 40, 72, 64, 2, False), (2, 8, 256, 256, 64, 1, False),
    (1, 8, 500, 500, 64, 1, False), (2, 16, 50, 50, 32, 1, False), (3, 4, 10, 10, 16, 1, False)])
def test_attention_forward(ops, engine, N, H, Tq, Tk, d, renorm, causal):
    from savqa_b200.functional import tc_attention_fits
    if engine == 0 and not tc_attention_fits(d, Tk):
        pytest.skip("tcgen05 engine takes d in {64,128} and Tk*d that fits one CTA's shared memory; others run on engine 1")
    q, k, v, graph, key_on, query_on = _attn_inputs(f"att/{N}/{H}/{Tq}/{Tk}/{d}", N, H, Tq, Tk, d, Tq == Tk)
    g_in = None if renorm == 0 else graph
    ref_out, ref_att = _attn_ref(q, k, v, g_in, key_on, query_on, N, H, Tq, Tk, d, causal, renorm)
    out, att = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(g_in), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, renorm, True, engine)
    torch.cuda.synchronize()
    # engine 1 is fp32 on the same bf16 inputs: summation order only.  engine 0 additionally rounds the un-normalised
    # probabilities to bf16 before P.V (one 2^-9 rounding) and uses ex2-based exp.
    assert rel(att, ref_att) < (1e-5 if engine == 1 else 2e-5), "attention probabilities"
    assert rel(out, ref_out) < (1e-5 if engine == 1 else 3e-3), "attention output"
    if Tq > 3 and renorm == 1:
        a4 = att.cpu().view(H, N, Tq, Tk)
        assert float(a4[:, 0, 1].abs().sum()) == 0.0 and float(a4[:, 0, 2].abs().sum()) == 0.0  # zero-edge rows stay exactly zero
    # broadcast graph [N,1,Tk] (decoder mask) must equal the expanded one
    if renorm == 1 and Tq > 1:
        g1 = graph[:, :1].contiguous()
        o1, _ = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(g1), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, 1, False, engine)
        o2, _ = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(g1.expand(N, Tq, Tk).contiguous()), dev(key_on), dev(query_on), N, H, Tq,
                                        Tk, d, causal, 1, False, engine)
        assert torch.equal(o1, o2)
    # the bit-packed form of the (0/1) graph must give bit-identical results and is itself bit exact
    if engine == 0 and renorm != 0 and Tq > 1:
        import fake_ops
        bits = ops.pack_graph_bits(dev(graph))
        assert torch.equal(bits.cpu(), fake_ops.pack_graph_bits(graph))
        ob, ab = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(graph), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, renorm, True,
                                         engine, graph_bits=bits)
        assert torch.equal(ob, out) and torch.equal(ab, att)
        b1 = ops.pack_graph_bits(dev(graph[:, :1].contiguous()))
        o3, _ = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(graph[:, :1].contiguous()), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal,
                                        renorm, False, engine, graph_bits=b1)
        o4, _ = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(graph[:, :1].contiguous()), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal,
                                        renorm, False, engine)
        assert torch.equal(o3, o4)


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("N,H,Tq,Tk,d,renorm,causal", [(3, 8, 56, 56, 64, 1, False), (2, 8, 1, 56, 64, 1, False), (2, 4, 36, 36, 16, 0, True),
                                                       (2, 8, 40, 72, 64, 2, False), (2, 4, 130, 130, 128, 1, False), (2, 16, 50, 50, 32, 1, False),
                                                       (2, 8, 128, 128, 64, 1, False), (2, 8, 100, 200, 64, 1, False), (2, 8, 36, 36, 64, 0, True),
                                                       (2, 4, 128, 128, 128, 1, False), (2, 8, 3, 256, 64, 1, False), (130, 8, 56, 56, 64, 1, False)])
def test_attention_backward(ops, engine, N, H, Tq, Tk, d, renorm, causal):
    import fake_ops
    if engine == 0 and not ops._bwd_one_cta_fits(d, Tq, Tk):
        pytest.skip("one CTA of the tcgen05 backward holds d in {64,128}, Tq <= 128, Tk <= 256; larger / 32-channel shapes are tiled on the "
                    "forward statistics (test_attention_tcgen05_small_heads_and_long_queries)")
    q, k, v, graph, key_on, query_on = _attn_inputs(f"attb/{N}/{H}/{Tq}/{Tk}/{d}", N, H, Tq, Tk, d, False)
    g_in = None if renorm == 0 else graph
    C = H * d
    dout = GS.randn(f"attb/{N}/{Tq}/{C}/dout", N * Tq, C)
    rq, rk, rv = torch.zeros(N * Tq, C, dtype=BF), torch.zeros(N * Tk, C, dtype=BF), torch.zeros(N * Tk, C, dtype=BF)
    fake_ops.graph_attention_bwd(q, k, v, g_in, key_on, query_on, N, H, Tq, Tk, d, causal, renorm, dout, rq, rk, rv)
    # gradients land in column slices of the fused [.., 3C] projection-gradient layout, as in functional.py
    dqkv = torch.zeros(N * max(Tq, Tk), 3 * C, dtype=BF, device="cuda")
    dq, dk, dv = dqkv[:N * Tq, :C], dqkv[:N * Tk, C:2 * C], dqkv[:N * Tk, 2 * C:]
    db = torch.zeros(3, C, device="cuda")
    ops.graph_attention_bwd(dev(q), dev(k), dev(v), dev(g_in), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, renorm, dev(dout), dq, dk, dv,
                            engine=engine, dbq=db[0], dbk=db[1], dbv=db[2])
    torch.cuda.synchronize()
    # bias gradients of the projections = column sums of the gated gradients the same kernel stored (fp32 sums of the
    # un-rounded values vs sums of the bf16-rounded stores: 2^-9 per element, averaged down over the rows)
    for b_, g_ in zip(db, (dq, dk, dv)):
        ref_b = g_.float().sum(0)
        assert float((b_ - ref_b).abs().max()) <= 1e-2 * float(ref_b.abs().max()) + 1e-6
    # engine 1: both sides round the fp32 gradients to bf16 once (2^-9 where the roundings differ).  engine 0 additionally
    # feeds bf16 dO, dS and W' to the tensor cores (three more 2^-9 roundings, fp32 accumulation).
    tol = 2e-3 if engine == 1 else 6e-3
    assert rel(dq, rq) < tol and rel(dk, rk) < tol and rel(dv, rv) < tol, (rel(dq, rq), rel(dk, rk), rel(dv, rv))
    if engine == 0 and g_in is not None:  # bit-packed graph: identical arithmetic, identical bits out
        d2 = torch.zeros_like(dqkv)
        ops.graph_attention_bwd(dev(q), dev(k), dev(v), dev(g_in), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, renorm, dev(dout),
                                d2[:N * Tq, :C], d2[:N * Tk, C:2 * C], d2[:N * Tk, 2 * C:], engine=engine, graph_bits=ops.pack_graph_bits(dev(g_in)))
        assert torch.equal(d2, dqkv)
    if engine == 0:
        # one-pass form: the forward kernel's row statistics {m, 1/Z, scale, beta} and output replace the recomputation of the
        # row maximum / sums, with t_i = <dO_i, O_i> instead of sum_j W_ij dW_ij (same quantity, different bf16 roundings)
        stats = torch.empty(H * N * Tq, 4, device="cuda")
        o_f, _ = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(g_in), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, renorm, False, 0,
                                         stats=stats)
        d3 = torch.zeros_like(dqkv)
        db3 = torch.zeros(3, C, device="cuda")
        ops.graph_attention_bwd(dev(q), dev(k), dev(v), dev(g_in), dev(key_on), dev(query_on), N, H, Tq, Tk, d, causal, renorm, dev(dout),
                                d3[:N * Tq, :C], d3[:N * Tk, C:2 * C], d3[:N * Tk, 2 * C:], engine=0, dbq=db3[0], dbk=db3[1], dbv=db3[2],
                                stats=stats, fwd_out=o_f)
        torch.cuda.synchronize()
        e3 = (rel(d3[:N * Tq, :C], rq), rel(d3[:N * Tk, C:2 * C], rk), rel(d3[:N * Tk, 2 * C:], rv))
        # t comes from the forward's bf16-rounded probabilities here, so sum_j dS_ij is zero only to ~2^-9 |t| instead of exactly:
        # dQ = dS K picks that residue up along the mean key (largest for a single query row, which the product path never
        # sends to this engine)
        assert max(e3) < (1e-2 if Tq < 8 else tol), e3


@pytest.mark.parametrize("N,H,Tq,Tk", [(3, 8, 128, 128), (2, 8, 100, 100), (2, 8, 72, 72), (2, 8, 40, 120), (130, 8, 128, 128),
                                       (3, 8, 56, 56), (2, 8, 64, 64), (2, 8, 36, 36), (2, 8, 20, 50)])
def test_attention_backward_shared_tile_kernel(ops, N, H, Tq, Tk):
    """The step's encoder backward (32 < Tk <= 128, bit-packed graph, forward statistics): the two-CTAs-per-SM kernel
    that shares one tile between W' and dS must agree with the fp32 restatement and -- bit for bit -- with the one-CTA kernel
    (same arithmetic, different schedule), masked keys and query rows included."""
    import os
    import fake_ops
    d, C = 64, H * 64
    q, k, v, graph, key_on, query_on = _attn_inputs(f"attb1/{N}/{H}/{Tq}/{Tk}", N, H, Tq, Tk, d, False)
    dout = GS.randn(f"attb1/{N}/{Tq}/{C}/dout", N * Tq, C)
    rq, rk, rv = torch.zeros(N * Tq, C, dtype=BF), torch.zeros(N * Tk, C, dtype=BF), torch.zeros(N * Tk, C, dtype=BF)
    fake_ops.graph_attention_bwd(q, k, v, graph, key_on, query_on, N, H, Tq, Tk, d, False, 1, dout, rq, rk, rv)
    bits = ops.pack_graph_bits(dev(graph))
    stats = torch.empty(H * N * Tq, 4, device="cuda")
    o_f, _ = ops.graph_attention_fwd(dev(q), dev(k), dev(v), dev(graph), dev(key_on), dev(query_on), N, H, Tq, Tk, d, False, 1, False, 0,
                                     graph_bits=bits, stats=stats)
    outs = []
    for off in (False, True):
        if off:
            os.environ["SAVQA_ATTN_BWD_ONE_TILE_OFF"] = "1"
        try:
            dqkv = torch.zeros(N * max(Tq, Tk), 3 * C, dtype=BF, device="cuda")
            db = torch.zeros(3, C, device="cuda")
            ops.graph_attention_bwd(dev(q), dev(k), dev(v), dev(graph), dev(key_on), dev(query_on), N, H, Tq, Tk, d, False, 1, dev(dout),
                                    dqkv[:N * Tq, :C], dqkv[:N * Tk, C:2 * C], dqkv[:N * Tk, 2 * C:], engine=0, dbq=db[0], dbk=db[1], dbv=db[2],
                                    graph_bits=bits, stats=stats, fwd_out=o_f)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("SAVQA_ATTN_BWD_ONE_TILE_OFF", None)
        outs.append((dqkv, db))
    (g1, b1), (g0, b0) = outs
    e = (rel(g1[:N * Tq, :C], rq), rel(g1[:N * Tk, C:2 * C], rk), rel(g1[:N * Tk, 2 * C:], rv))
    assert max(e) < 6e-3, e
    assert torch.equal(g1, g0)
    for x, y in zip(b1, b0):  # bias gradients: same addends, atomics in a different order
        assert float((x - y).abs().max()) <= 1e-5 * float(y.abs().max()) + 1e-6


# ------------------------------------------------------------------------------------------------ loss / adam
def test_answer_loss_and_adam(ops):
    B, ncls = 37, 1845
    L = [GS.randn(f"loss/{i}", B, ncls) for i in range(3)]
    ans = GS.randint("loss/ans", 0, ncls, B)
    Lr = [x.clone().requires_grad_(True) for x in L]
    ref = O.answer_loss(*Lr, ans)
    ref.backward()
    loss, grads = ops.answer_loss(dev(L[0]), dev(L[1]), dev(L[2]), dev(ans), 0.1, 1.0, True)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    for g_, r_ in zip(grads, Lr):
        assert rel(g_, r_.grad) < 1e-5
    p = GS.randn("adam/p", 10000)
    g = GS.randn("adam/g", 10000)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    pd, m, v = dev(p.clone()), torch.zeros(10000, device="cuda"), torch.zeros(10000, device="cuda")
    mirror = torch.zeros(10000, device="cuda", dtype=BF)
    for step in (1, 2, 3):
        pr.grad = g.clone() * step
        opt.step()
        ops.adam_step(pd, dev(g * step), m, v, 1e-3, 0.9, 0.999, 1e-8, step, param_bf16=mirror)
    assert rel(pd, pr.detach()) < 1e-6
    assert torch.equal(mirror, pd.to(BF))  # the bf16 mirror the GEMMs read
    # unaligned views take the scalar path
    pd2, m2, v2 = dev(p.clone())[1:9998], torch.zeros(9999, device="cuda")[1:9998], torch.zeros(9999, device="cuda")[1:9998]
    pr2 = p[1:9998].clone().requires_grad_(True)
    opt2 = torch.optim.Adam([pr2], lr=1e-3)
    pr2.grad = g[1:9998].clone()
    opt2.step()
    ops.adam_step(pd2, dev(g)[1:9998], m2, v2, 1e-3, 0.9, 0.999, 1e-8, 1)
    assert rel(pd2, pr2.detach()) < 1e-6


def test_deferred_row_adam_equals_dense_adam(ops):
    """Row-sparse word-table optimizer == the reference's dense torch.optim.Adam (rows replay the zero-gradient steps they
    missed): tests/parity_cases.deferred_adam_case."""
    import parity_cases as PC
    for seed in range(3):
        PC.deferred_adam_case(ops, "cuda", steps=15, seed=seed)
    # the word rows' width (three 16-byte groups per lane, the last one ragged), a width the 16-byte kernel does not take (scalar
    # kernel), one that needs all three groups
    for width in (300, 10, 384):
        PC.deferred_adam_case(ops, "cuda", steps=12, rows=24, width=width, seed=7)
    # a long absence takes the closed-form branch (> 256 missed steps): the parameter update of those steps is below fp32 resolution
    rows, width = 4, 8
    p = torch.randn(rows, width, device="cuda")
    ref = p.clone().cpu().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2)
    grad, m, v = (torch.zeros(rows, width, device="cuda") for _ in range(3))
    stamp = torch.zeros(rows, dtype=torch.int32, device="cuda")
    g0 = torch.randn(rows, width)
    idx = torch.arange(rows, device="cuda")
    for step in (1, 400):
        gd = g0 if step == 1 else 2 * g0
        while opt.state and int(opt.state[ref]["step"]) < step - 1:
            ref.grad = torch.zeros_like(ref)
            opt.step()
        ref.grad = gd.clone()
        opt.step()
        ops.scatter_add_rows(grad, idx, gd.cuda())
        ops.adam_rows(p, grad, m, v, stamp, idx, 1e-2, 0.9, 0.999, 1e-8, step, apply=True)
    assert float((p.cpu() - ref.detach()).abs().max()) < 1e-5


# ------------------------------------------------------------------------------------------------ cluster GEMM + row-wise epilogue
@pytest.mark.parametrize("M,N,K,b_mn", [(128, 512, 512, False), (128, 512, 2048, False), (100, 512, 512, True), (128, 256, 192, False),
                                        (77, 64, 64, False), (256, 512, 512, True), (300, 128, 2048, True)])
def test_gemm_rowln_layernorm_modes(ops, M, N, K, b_mn):
    """savqa_gemm_rowln modes 1 (Linear -> mask -> residual -> LayerNorm) and 2 (dgrad -> residual -> LayerNorm backward -> ReLU gate)
    against the fp32 restatement of the same formulas (tests/fake_ops.py), incl. ragged row counts, clusters of 1 / 2 / 4 / 8 CTAs,
    several 128-row blocks, K-major and MN-major weights."""
    import fake_ops as F
    g = torch.Generator().manual_seed(M * 7 + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(BF)
    w = (torch.randn(K, N, generator=g) if b_mn else torch.randn(N, K, generator=g)).mul(K ** -0.5).to(BF)
    bias, rowscale = torch.randn(N, generator=g) * 0.1, (torch.rand(M, generator=g) < 0.8).float()
    res, gamma, beta = torch.randn(M, N, generator=g), torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1
    res[3] = 0
    rowscale[3] = 0  # an all-zero LayerNorm input row when the bias is masked too: exercised below through rowscale = 0, res = 0
    outs = lambda dev: dict(act_bf16=torch.zeros(M, N, dtype=BF, device=dev), pre=torch.zeros(M, N, device=dev), y=torch.zeros(M, N, device=dev),  # noqa: E731
                            y_bf16=torch.zeros(M, N, dtype=BF, device=dev), on=torch.zeros(M, device=dev), stats=torch.zeros(M, 2, device=dev))
    ref, got = outs("cpu"), outs("cuda")
    F.gemm_rowln(a, w, M, N, K, 1, b_mn=b_mn, bias=bias, relu=True, rowscale=rowscale, res=res, gamma=gamma, beta=beta, eps=1e-8, **ref)
    ops.gemm_rowln(dev(a), dev(w), M, N, K, 1, b_mn=b_mn, bias=dev(bias), relu=True, rowscale=dev(rowscale), res=dev(res), gamma=dev(gamma),
                   beta=dev(beta), eps=1e-8, **got)
    assert torch.equal(got["act_bf16"].cpu(), ref["act_bf16"]) or rel(got["act_bf16"].float(), ref["act_bf16"].float()) < 2e-3
    assert rel(got["pre"], ref["pre"]) < 2e-3 and rel(got["y"], ref["y"]) < 3e-3 and rel(got["stats"], ref["stats"]) < 2e-3
    assert torch.equal(got["on"].cpu(), ref["on"])
    assert torch.equal(got["y"][3].cpu(), beta)  # constant (zero) LayerNorm input row -> exactly beta
    assert rel(got["y_bf16"].float(), got["y"].to(BF).float()) < 1e-6
    # mode 2 on the kernel's own forward outputs
    dy_a = (torch.randn(M, K, generator=g) * 0.3).to(BF)
    res2 = torch.randn(M, N, generator=g)
    gate = (torch.randn(M, N, generator=g)).to(BF)
    pre_d, stats_d = got["pre"].clone(), got["stats"].clone()
    o2 = lambda dev: dict(y=torch.zeros(M, N, device=dev), y_bf16=torch.zeros(M, N, dtype=BF, device=dev), dxg_bf16=torch.zeros(M, N, dtype=BF, device=dev),  # noqa: E731
                          dgamma=torch.zeros(N, device=dev), dbeta=torch.zeros(N, device=dev), dxsum=torch.zeros(N, device=dev))
    r2, g2 = o2("cpu"), o2("cuda")
    F.gemm_rowln(dy_a, w, M, N, K, 2, b_mn=b_mn, res=res2, pre=pre_d.cpu(), stats=stats_d.cpu(), gamma=gamma, eps=1e-8, gate=gate,
                 rowscale=rowscale, **r2)
    ops.gemm_rowln(dev(dy_a), dev(w), M, N, K, 2, b_mn=b_mn, res=dev(res2), pre=pre_d, stats=stats_d, gamma=dev(gamma), eps=1e-8, gate=dev(gate),
                   rowscale=dev(rowscale), **g2)
    live = torch.ones(M, dtype=torch.bool)
    live[3] = False  # the sigma == 0 row carries 1 / eps = 1e8-scaled values: compared apart
    assert rel(g2["y"].cpu()[live], r2["y"][live]) < 2e-3
    assert rel(g2["y"].cpu()[3], r2["y"][3]) < 2e-3
    assert rel(g2["dxg_bf16"].float().cpu()[live], r2["dxg_bf16"].float()[live]) < 6e-3
    for k_ in ("dbeta", "dgamma"):
        assert rel(g2[k_], r2[k_]) < 2e-3, k_
    # plain mode with a ReLU gate and a residual (the feedforward's dgrad)
    yp_r, yp_g = torch.zeros(M, N, dtype=BF), torch.zeros(M, N, dtype=BF, device="cuda")
    F.gemm_rowln(dy_a, w, M, N, K, 0, b_mn=b_mn, gate=gate, y_bf16=yp_r)
    ops.gemm_rowln(dev(dy_a), dev(w), M, N, K, 0, b_mn=b_mn, gate=dev(gate), y_bf16=yp_g)
    assert rel(yp_g.float(), yp_r.float()) < 3e-3


def test_gemm_rowln_plain_wide(ops):
    """Plain mode at the decoder's conv1 shape (N = 2048: 32 independent CTAs) with bias + ReLU, fp32 and bf16 outputs."""
    import fake_ops as F
    M, N, K = 128, 2048, 512
    g = torch.Generator().manual_seed(5)
    a, w, bias = (torch.randn(M, K, generator=g) * 0.5).to(BF), (torch.randn(N, K, generator=g) * K ** -0.5).to(BF), torch.randn(N, generator=g) * 0.1
    yr, ybr = torch.zeros(M, N), torch.zeros(M, N, dtype=BF)
    yg, ybg = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, dtype=BF, device="cuda")
    F.gemm_rowln(a, w, M, N, K, 0, bias=bias, relu=True, y=yr, y_bf16=ybr)
    ops.gemm_rowln(dev(a), dev(w), M, N, K, 0, bias=dev(bias), relu=True, y=yg, y_bf16=ybg)
    assert rel(yg, yr) < 2e-3 and rel(ybg.float(), ybr.float()) < 3e-3


def test_layernorm_forward_statistics(ops):
    """The optional {mean, sigma} output of savqa_residual_layernorm_fwd (what the fused dgrad + LayerNorm-backward epilogue reads)."""
    for rows, C in ((37, 512), (19, 64)):
        x, res = GS.randn(f"lnst/x{C}", rows, C), GS.randn(f"lnst/r{C}", rows, C)
        gamma, beta = GS.rand(f"lnst/g{C}", C, lo=0.8, hi=1.2), GS.randn(f"lnst/b{C}", C)
        stats = torch.zeros(rows, 2, device="cuda")
        ops.layernorm_fwd(dev(x), dev(res), dev(gamma), dev(beta), 1e-8, True, True, True, stats=stats)
        pre = x + res
        assert rel(stats[:, 0], pre.mean(-1)) < 1e-5 and rel(stats[:, 1], pre.std(-1)) < 1e-5


@pytest.mark.parametrize("N,H,T,d", [(3, 16, 40, 32), (2, 16, 100, 32), (2, 8, 200, 64), (2, 8, 256, 64), (2, 16, 160, 32), (2, 16, 256, 32),
                                     (2, 8, 299, 64), (2, 4, 200, 128)])
def test_attention_tcgen05_small_heads_and_long_queries(ops, N, H, T, d):
    """BASELINE configs[4] shapes that used to fall to the CUDA-core engine: 32-channel heads (16 heads x 512: run as zero-padded
    64-wide tiles, score scale 1/sqrt(32)) and the backward pass for 128 < Tq <= 256 (two query tiles).  Forward and backward of the
    tcgen05 engine against the fp32 CUDA-core verification engine (itself checked against the restatement above)."""
    from savqa_b200 import _lib
    C = H * d
    g = torch.Generator().manual_seed(N * 1000 + T + d)
    qkv = (torch.randn(N * T, 3 * C, generator=g) * 0.5).relu().to(BF).cuda()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    graph = (torch.rand(N, T, T, generator=g) < 0.3).float().cuda()
    graph[:, torch.arange(T), torch.arange(T)] = 1
    graph[:, 1] = 0
    key_on = torch.ones(N * T).cuda()
    key_on[T - 2:T] = 0
    q_on = torch.ones(N * T).cuda()
    q_on[3] = 0
    bits = ops.pack_graph_bits(graph)
    stats = torch.empty(H * N * T, 4, device="cuda")
    c0 = _lib.launch_counts()
    o0, _ = ops.graph_attention_fwd(q, k, v, graph, key_on, q_on, N, H, T, T, d, False, 1, False, 0, graph_bits=bits, stats=stats)
    o1, _ = ops.graph_attention_fwd(q, k, v, graph, key_on, q_on, N, H, T, T, d, False, 1, False, 1)
    assert rel(o0, o1) < 3e-3, rel(o0, o1)
    dout = torch.randn(N * T, C, generator=g).cuda()
    outs = []
    for eng in (0, 1):
        dqkv = torch.zeros(N * T, 3 * C, device="cuda", dtype=BF)
        db = [torch.zeros(C, device="cuda") for _ in range(3)]
        ops.graph_attention_bwd(q, k, v, graph, key_on, q_on, N, H, T, T, d, False, 1, dout, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:],
                                engine=eng, dbq=db[0], dbk=db[1], dbv=db[2], graph_bits=bits if eng == 0 else None,
                                stats=stats if eng == 0 else None, fwd_out=o0 if eng == 0 else None)
        outs.append((dqkv.float(), db))
    c1 = _lib.launch_counts()
    assert c1["attn_fwd_tc"] > c0["attn_fwd_tc"] and (c1["attn_bwd_tc"] + c1["attn_bwd_tc_shared"]) > (c0["attn_bwd_tc"] + c0["attn_bwd_tc_shared"])
    assert rel(outs[0][0], outs[1][0]) < 1.5e-2, rel(outs[0][0], outs[1][0])
    for a_, b_ in zip(outs[0][1], outs[1][1]):
        assert rel(a_, b_) < 1.5e-2


def test_two_host_threads_share_the_library(ops):
    """Two host threads launching through the C ABI at the same time (own streams, different per-thread SM limits, the shared
    TMA-descriptor cache behind its mutex): results identical to the single-threaded ones."""
    import threading
    M, N, K = 1024, 512, 512
    g = torch.Generator().manual_seed(0)
    a = (torch.randn(M, K, generator=g) * 0.5).to(BF).cuda()
    ws = [(torch.randn(N, K, generator=g) * K ** -0.5).to(BF).cuda() for _ in range(2)]
    ref = []
    for w in ws:
        o = torch.empty(M, N, device="cuda")
        ops.gemm(a, w, M, N, K, out_f32=o)
        ref.append(o.clone())
    torch.cuda.synchronize()
    outs, errs = [None, None], []

    def work(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st), ops.gemm_sm_limit(32 if i == 0 else 100):
                for _ in range(40):
                    o = torch.empty(M, N, device="cuda")
                    ops.gemm(a, ws[i], M, N, K, out_f32=o)
                    x = torch.randn(64 + i, 512, device="cuda")
                    ops.layernorm_fwd(x, None, torch.ones(512, device="cuda"), torch.zeros(512, device="cuda"), 1e-8, False, False, False)
                outs[i] = o
            st.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for o, r in zip(outs, ref):
        assert torch.equal(o, r)
