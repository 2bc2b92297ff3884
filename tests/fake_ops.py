"""CPU stand-ins for savqa_b200.ops, used ONLY by the `not gpu` host-logic tests (tests/test_host_logic.py).

The product never imports this file.  Each function restates, with torch CPU ops, the contract of the matching
C-ABI kernel (including the explicit backward formulas the CUDA kernels implement -- NOT autograd), so that the
autograd wiring in savqa_b200/functional.py and the module composition can be checked against the oracle on a box
without a GPU.  The kernels themselves are checked on the B200 by the `gpu` tests.
"""
from __future__ import annotations

import contextlib

import torch

BF16, F32 = torch.bfloat16, torch.float32
MASK_FILL = -4294967296.0


def pad8(n):
    return (n + 7) // 8 * 8


def build_masks(first_mask, q_mask, q_graph, first_graph, dec_mask_on):
    B, V, Q = first_mask.shape[0], first_mask.shape[1], q_mask.shape[1]
    T = V + Q
    gd = torch.zeros(B, T, T)
    gd[:, V:, V:] = q_mask.float()
    g = torch.ones(B, T, T)
    if first_graph is not None:
        g[:, :V, :V] = first_graph.float()
    g[:, V:, V:] = q_graph.float()
    dm = torch.zeros(B, 1, T)
    if dec_mask_on:
        dm[:, 0, :V] = (first_mask.float().sum(-1) != 0).float()
        dm[:, 0, V:] = (q_mask.float().sum(-1) != 0).float()
    attach_graph_bits(gd)
    attach_graph_bits(g)
    return gd, g, dm


def build_masks_compact(first_len, q_len, q_graph_bits, first_graph_bits, V, Q, dec_mask_on):
    from savqa_b200 import collate

    def block(n, N):
        valid = torch.arange(N)[None, :] < n[:, None].long()
        return (valid[:, :, None] & valid[:, None, :]).to(torch.int32)
    fg = collate.unpack_adjacency(first_graph_bits, V) if first_graph_bits is not None else None
    return build_masks(block(first_len, V), block(q_len, Q), collate.unpack_adjacency(q_graph_bits, Q), fg, dec_mask_on)


def mil_nce_fwd(pn_h, vis_h, mask, loc, nodes, B, V, M, topN, h):
    """savqa_mil_nce_fwd restated (AttModel_x3.py:365-379): per-object scores, mil_nce_obj, softmax-refined object rows into nodes."""
    n = B * V * topN
    pos, neg, vis = pn_h[:n, :h].float().view(B * V, topN, h), pn_h[n:, :h].float().view(B * V, topN, h), vis_h[:, :h].float()
    rp, rn = (pos * vis[:, None]).sum(-1), (neg * vis[:, None]).sum(-1)
    sn = (mask.view(B * V, topN).float() * rn).clamp(min=1e-6)
    term = (1e-6 + torch.log(torch.tensor(float(topN)))) - torch.logsumexp(sn, 1)
    obj = (term.sum() / (2.0 * B * V)).reshape(1)
    refined = (torch.softmax(rp, 1)[:, :, None] * pos).sum(1)
    lo = loc.view(B * V)
    ok = (lo >= 0) & (lo < M)
    rows = (torch.arange(B * V) // V) * M + lo
    nodes[rows[ok], :h] = refined[ok].to(BF16)
    return torch.stack([rp.reshape(-1), rn.reshape(-1)]), obj


def mil_nce_bwd(pn_h, vis_h, mask, loc, raw, d_nodes, d_obj, B, V, M, topN, h):
    """savqa_mil_nce_bwd restated with the kernel's explicit formulas (not autograd)."""
    n = B * V * topN
    pos, neg, vis = pn_h[:n, :h].float().view(B * V, topN, h), pn_h[n:, :h].float().view(B * V, topN, h), vis_h[:, :h].float()
    rp, rn = raw[0].view(B * V, topN), raw[1].view(B * V, topN)
    mk = mask.view(B * V, topN).float()
    g = float(d_obj) / (2.0 * B * V) if d_obj is not None else 0.0
    mrn = mk * rn
    q = torch.softmax(mrn.clamp(min=1e-6), 1)
    dn = -g * q * torch.where(mrn >= 1e-6, mk, torch.zeros_like(mk))
    p = torch.softmax(rp, 1)
    lo = loc.view(B * V)
    ok = (lo >= 0) & (lo < M)
    dr = torch.zeros(B * V, h)
    if d_nodes is not None:
        rows = (torch.arange(B * V) // V) * M + lo
        dr[ok] = d_nodes[rows[ok], :h]
    dp = (dr[:, None] * pos).sum(-1)
    draw = p * (dp - (p * dp).sum(1, keepdim=True))
    dpos = (p[:, :, None] * dr[:, None] + draw[:, :, None] * vis[:, None]) * (pos > 0)
    dneg = (dn[:, :, None] * vis[:, None]) * (neg > 0)
    dvis = ((draw[:, :, None] * pos).sum(1) + (dn[:, :, None] * neg).sum(1)) * (vis > 0)
    return torch.cat([dpos.reshape(n, h), dneg.reshape(n, h)]).to(BF16), dvis.to(BF16)


def gather_rows(table, idx, scale=1.0, want_f32=True, want_bf16=False):
    out = table[idx.reshape(-1)]
    if scale != 1.0:
        out = out * scale
    o16 = None
    if want_bf16:
        o16 = torch.zeros(out.shape[0], pad8(out.shape[1]), dtype=BF16)
        o16[:, :out.shape[1]] = out.to(BF16)
    return (out if want_f32 else None), o16


def scatter_add_rows(dtable, idx, dout, scale=1.0, skip_row=-1):
    idx = idx.reshape(-1)
    keep = idx != skip_row
    add = dout.reshape(idx.numel(), -1)[keep] * scale
    if dtable.dtype == torch.int64:  # the Q15.48 fixed-point accumulator (savqa_scatter_add_rows_q48)
        add = torch.round(add.double() * float(2 ** 48)).to(torch.int64)
    dtable.index_add_(0, idx[keep], add)


def cast_bf16(src, out=None, pad_to=None):
    rows, cols = src.reshape(-1, src.shape[-1]).shape
    if out is None:
        pad_to = pad8(cols) if pad_to is None else pad_to
        out = torch.zeros(rows, pad_to, dtype=BF16)
    out[:, :cols] = src.reshape(rows, cols).to(BF16)
    if pad_to is not None and pad_to > cols:
        out[:, cols:pad_to] = 0
    return out


def cast_transpose_bf16(src, out, pad_to=None):
    rows, cols = src.shape
    out[:, :rows] = src.t().to(BF16)
    return out


def row_nonzero(x, want_bf16=True):
    x2 = x.reshape(-1, x.shape[-1])
    on = (x2.sum(-1) != 0).float()
    xb = None
    if want_bf16:
        xb = torch.zeros(x2.shape[0], pad8(x2.shape[1]), dtype=BF16)
        xb[:, :x2.shape[1]] = x2.to(BF16)
    return on, xb


def fill_zero(t, max_blocks=0):
    t.zero_()


def relu_gate_bf16(dy, act, group_rows=0, group_stride=0):
    if group_rows > 0:
        rows = act.shape[0]
        r = torch.arange(rows)
        dy = dy[(r // group_rows) * group_stride + (r % group_rows)]
    rows, cols = dy.shape
    out = torch.zeros(rows, pad8(cols), dtype=BF16)
    out[:, :cols] = torch.where(act[:, :cols].float() > 0, dy.float(), torch.zeros(())).to(BF16)
    return out


def colsum_bf16(x, out):
    out += x.float().sum(0)[: out.numel()]


def layernorm_fwd(x, res, gamma, beta, eps, save_pre, want_bf16, want_on, stats=None):
    pre = x if res is None else x + res
    C = pre.shape[-1]
    mean = pre.mean(-1, keepdim=True)
    c = pre - mean
    sigma = torch.sqrt((c * c).sum(-1, keepdim=True) / (C - 1))
    y = gamma * c / (sigma + eps) + beta
    if stats is not None:
        stats.view(-1, 2)[:mean.numel()] = torch.cat([mean.reshape(-1, 1), sigma.reshape(-1, 1)], 1)
    return y, (pre.clone() if save_pre else None), (y.to(BF16) if want_bf16 else None), ((y.reshape(-1, C).sum(-1) != 0).float() if want_on else None)


def layernorm_bwd(dy, pre, gamma, eps, dgamma, dbeta, dres_in=None, want_bf16=False, dxsum=None):
    # the explicit formula of csrc/layernorm.cu
    C = pre.shape[-1]
    mean = pre.mean(-1, keepdim=True)
    c = pre - mean
    sigma = torch.sqrt((c * c).sum(-1, keepdim=True) / (C - 1))
    s = sigma + eps
    g = dy * gamma
    mg = g.mean(-1, keepdim=True)
    dot = (g * c).sum(-1, keepdim=True)
    k2 = torch.where(sigma > 0, dot / ((C - 1) * sigma * s * s), torch.zeros(()))
    dx = (g - mg) / s - c * k2
    if dres_in is not None:
        dx = dx + dres_in
    if dgamma is not None:
        dgamma += (dy * c / s).reshape(-1, C).sum(0)
    if dbeta is not None:
        dbeta += dy.reshape(-1, C).sum(0)
    if dxsum is not None:
        dxsum += dx.reshape(-1, C).sum(0)
    return dx, (dx.to(BF16) if want_bf16 else None)


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, bias=None, res=None, rowtab=None, rowtab_period=0, gate=None, relu=False,
         alpha=1.0, out_f32=None, out_bf16=None, accumulate=0, split_k=1, colsum=None):
    assert a.dtype == BF16 and b.dtype == BF16
    A = a[:K, :M].float().t() if a_mn else a[:M, :K].float()
    Bm = b[:K, :N].float() if b_mn else b[:N, :K].float().t()
    v = alpha * (A @ Bm)
    if bias is not None:
        v = v + bias[:N]
    if res is not None:
        v = v + res[:M, :N]
    if rowtab is not None:
        v = v + rowtab[:rowtab_period, :N].repeat((M + rowtab_period - 1) // rowtab_period, 1)[:M]
    if relu:
        v = torch.relu(v)
    if gate is not None:
        v = v * (gate[:M, :N].float() > 0)
    if out_f32 is not None:
        if accumulate:
            out_f32[:M, :N] += v
        else:
            out_f32[:M, :N] = v
    if out_bf16 is not None:
        out_bf16[:M, :N] = v.to(BF16)
    if colsum is not None:
        colsum[:N] += v.sum(0)


def rowln_fits(M, N, K):
    return N in (64, 128, 256, 512) and K % 8 == 0 and M >= 1


def gemm_rowln(a, b, M, N, K, mode, *, b_mn=False, bias=None, relu=False, rowscale=None, res=None, gate=None, gamma=None, beta=None,
               eps=1e-8, act_bf16=None, pre=None, y=None, y_bf16=None, on=None, stats=None, dxg_bf16=None, dgamma=None, dbeta=None,
               dxsum=None):
    """savqa_gemm_rowln restated: bf16 operands, fp32 accumulate, the three epilogue modes with the kernel's explicit formulas."""
    A = a[:M, :K].float()
    W = b[:K, :N].float().t() if b_mn else b[:N, :K].float()
    acc = A @ W.t()
    if mode != 2:
        v = acc + (bias[:N] if bias is not None else 0.0)
        if relu:
            v = v.clamp(min=0)
        if gate is not None:
            v = torch.where(gate[:M, :N].float() > 0, v, torch.zeros_like(v))
        if act_bf16 is not None:
            act_bf16[:M, :N] = v.to(BF16)
            v = act_bf16[:M, :N].float()
        if mode == 0:
            if res is not None:
                v = v + res[:M, :N]
            if y is not None:
                y[:M, :N] = v
            if y_bf16 is not None:
                y_bf16[:M, :N] = v.to(BF16)
            return
        if rowscale is not None:
            v = v * rowscale[:M, None]
        if res is not None:
            v = v + res[:M, :N]
        if pre is not None:
            pre[:M, :N] = v
        mean = v.mean(-1, keepdim=True)
        sigma = ((v - mean) ** 2).sum(-1, keepdim=True).div(N - 1).sqrt()
        out = gamma[:N] * (v - mean) / (sigma + eps) + beta[:N]
        if y is not None:
            y[:M, :N] = out
        if y_bf16 is not None:
            y_bf16[:M, :N] = out.to(BF16)
        if on is not None:
            on[:M] = (out.sum(-1) != 0).float()
        if stats is not None:
            stats.view(-1, 2)[:M] = torch.cat([mean, sigma], 1)
        return
    dy = acc + (res[:M, :N] if res is not None else 0.0)
    st = stats.view(-1, 2)[:M]
    mean, sigma = st[:, :1], st[:, 1:2]
    c = pre[:M, :N] - mean
    s_ = sigma + eps
    if dbeta is not None:
        dbeta[:N] += dy.sum(0)
    if dgamma is not None:
        dgamma[:N] += (dy * c / s_).sum(0)
    g = dy * gamma[:N]
    k2 = torch.where(sigma > 0, (g * c).sum(-1, keepdim=True) / ((N - 1) * sigma * s_ * s_), torch.zeros_like(sigma))
    dx = (g - g.mean(-1, keepdim=True)) / s_ - c * k2
    if dxsum is not None:
        dxsum[:N] += dx.sum(0)
    if y is not None:
        y[:M, :N] = dx
    if y_bf16 is not None:
        y_bf16[:M, :N] = dx.to(BF16)
    if dxg_bf16 is not None:
        t = dx * (rowscale[:M, None] if rowscale is not None else 1.0)
        dxg_bf16[:M, :N] = torch.where(gate[:M, :N].float() > 0, t, torch.zeros_like(t)).to(BF16)


def tc_attention_bwd_fits(d, Tq, Tk):
    return False


def gemm_grouped(problems, N, *, a_mn=False, b_mn=False, split_k=1):
    for p in problems:
        kw = {k: v for k, v in p.items() if k not in ("a", "b", "M", "K")}
        gemm(p["a"], p["b"], p["M"], N, p["K"], a_mn=a_mn, b_mn=b_mn, **kw)


def wgrad_grouped(items):
    for dy, x, n_out, k_in, out in items:
        wgrad(dy, x, n_out, k_in, out)


def split_k_for(tiles, k_blocks):
    return 1


@contextlib.contextmanager
def gemm_sm_limit(sms):
    yield


def wgrad(dy, x, n_out, k_in, out):
    gemm(dy, x, n_out, k_in, dy.shape[0], a_mn=True, b_mn=True, out_f32=out, accumulate=2)


def _heads(t, N, T, H, d):
    return t[:, : H * d].float().reshape(N, T, H, d).permute(2, 0, 1, 3)  # [H,N,T,d]


def _weights(q, k, graph, key_on, N, H, Tq, Tk, d, causal, renorm):
    Q, K = _heads(q, N, Tq, H, d), _heads(k, N, Tk, H, d)
    S = torch.matmul(Q, K.transpose(-1, -2)) / (d ** 0.5)  # [H,N,Tq,Tk]
    fixed = (key_on.reshape(1, N, 1, Tk) == 0).expand_as(S).clone()
    if causal:
        fixed |= ~torch.ones(Tq, Tk, dtype=torch.bool).tril().reshape(1, 1, Tq, Tk)
    S = torch.where(fixed, torch.full((), MASK_FILL), S)
    P = torch.softmax(S, -1)
    if renorm == 0:
        return P, P, fixed, None, None
    G = graph.reshape(1, N, graph.shape[1], Tk).expand(H, N, Tq, Tk)
    A = G * P
    r = A.abs().sum(-1, keepdim=True)
    if renorm == 1:
        W = A / r.clamp_min(1e-12)
    else:
        W = A / (A.sum(-1, keepdim=True) + 1e-7)
    return P, W, fixed, r, G


def pack_graph_bits(graph):
    N, Tq, Tk = graph.shape
    wpr = (Tk + 31) // 32
    padded = torch.zeros(N, Tq, wpr * 32, dtype=torch.int64)
    padded[:, :, :Tk] = (graph != 0).long()
    words = (padded.reshape(N, Tq, wpr, 32) << torch.arange(32)).sum(-1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


def attach_graph_bits(graph):
    graph._savqa_bits = (pack_graph_bits(graph), graph._version, graph.data_ptr())
    return graph


def graph_bits_of(graph):
    tag = getattr(graph, "_savqa_bits", None) if graph is not None else None
    if tag is not None and tag[1] == graph._version and tag[2] == graph.data_ptr():
        return tag[0]
    return None


def _check_bits(graph, graph_bits):
    if graph_bits is not None:  # what the kernels would read instead of the fp32 graph must say the same thing
        assert graph is not None and torch.equal(graph_bits, pack_graph_bits(graph.float()))
        assert bool(((graph == 0) | (graph == 1)).all())


def graph_attention_fwd(q, k, v, graph, key_on, query_on, N, H, Tq, Tk, d, causal, renorm, want_att, engine, graph_bits=None, stats=None):
    _check_bits(graph, graph_bits)
    P, W, _, _, _ = _weights(q, k, graph, key_on, N, H, Tq, Tk, d, causal, renorm)
    V = _heads(v, N, Tk, H, d)
    Wq = W * query_on.reshape(1, N, Tq, 1)
    O = torch.matmul(Wq, V)  # [H,N,Tq,d]
    out = O.permute(1, 2, 0, 3).reshape(N * Tq, H * d).contiguous()
    att = W.reshape(H * N, Tq, Tk).contiguous() if want_att else None
    return out, att


def graph_attention_bwd(q, k, v, graph, key_on, query_on, N, H, Tq, Tk, d, causal, renorm, dout, dq, dk, dv, engine=None, dbq=None,
                        dbk=None, dbv=None, graph_bits=None, stats=None, fwd_out=None):
    _check_bits(graph, graph_bits)
    # the explicit formulas of csrc/attn_simt.cu (attn_bwd_rows_kernel / attn_bwd_keys_kernel)
    P, W, fixed, r, G = _weights(q, k, graph, key_on, N, H, Tq, Tk, d, causal, renorm)
    Q, K, V = _heads(q, N, Tq, H, d), _heads(k, N, Tk, H, d), _heads(v, N, Tk, H, d)
    dO = dout[:, : H * d].reshape(N, Tq, H, d).permute(2, 0, 1, 3)
    qon = query_on.reshape(1, N, Tq, 1)
    dW = torch.matmul(dO, V.transpose(-1, -2)) * qon
    t = (W * dW).sum(-1, keepdim=True)
    if renorm == 1:
        clamped = r < 1e-12
        dS = torch.where(clamped, W * dW - P * t, W * (dW - t))
    elif renorm == 2:
        dS = W * (dW - t) - P * t * (1 - W.sum(-1, keepdim=True))
    else:
        dS = W * (dW - t)
    dS = torch.where(fixed, torch.zeros(()), dS) / (d ** 0.5)
    dQ = torch.matmul(dS, K) * (Q > 0)
    dK = torch.matmul(dS.transpose(-1, -2), Q) * (K > 0)
    dV = torch.matmul((W * qon).transpose(-1, -2), dO) * (V > 0)
    dq[:, : H * d] = dQ.permute(1, 2, 0, 3).reshape(N * Tq, H * d).to(BF16)
    dk[:, : H * d] = dK.permute(1, 2, 0, 3).reshape(N * Tk, H * d).to(BF16)
    dv[:, : H * d] = dV.permute(1, 2, 0, 3).reshape(N * Tk, H * d).to(BF16)
    for db, g, T in ((dbq, dQ, Tq), (dbk, dK, Tk), (dbv, dV, Tk)):
        if db is not None:
            db += g.permute(1, 2, 0, 3).reshape(N * T, H * d).sum(0)


def answer_loss(lc, lv, ls, answer, epsilon, grad_scale, want_grads):
    B, ncls = lc.shape
    t = torch.full((B, ncls), epsilon / ncls)
    t[torch.arange(B), answer] += 1 - epsilon
    loss = torch.zeros(1)
    grads = []
    for L in (lc, lv, ls):
        lsm = torch.log_softmax(L, -1)
        loss -= (t * lsm).sum() / (3 * B)
        grads.append(grad_scale * (lsm.exp() - t) / (3 * B))
    return loss, (tuple(grads) if want_grads else None)


def adam_advance(dyn, lr, beta1, beta2):
    s = float(dyn[2]) + 1.0
    dyn[0] = lr / (1.0 - beta1 ** s)
    dyn[1] = (1.0 - beta2 ** s) ** 0.5
    dyn[2] = s


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, dyn=None, param_bf16=None):
    if dyn is not None:
        step = int(dyn[2])  # the device-resident step counter wins (savqa_adam_advance)
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    param.sub_((lr / bc1) * exp_avg / (exp_avg_sq.sqrt() / (bc2 ** 0.5) + eps))
    if param_bf16 is not None:
        param_bf16.copy_(param.to(BF16))


def adam_rows(param, grad, exp_avg, exp_avg_sq, row_stamp, idx, lr, beta1, beta2, eps, step, dyn=None, apply=True):
    """Deferred Adam: replay the zero-gradient steps a row missed (dense torch.optim.Adam semantics), then this step's update."""
    if dyn is not None:
        step = int(dyn[2])
    target = step if apply else step - 1
    if target < 1:
        return
    rows = torch.unique(idx.reshape(-1)) if idx is not None else torch.arange(param.shape[0])
    rows = rows[(rows >= 0) & (rows < param.shape[0])]
    for r in rows.tolist():
        old = int(row_stamp[r])
        if old >= target:
            continue
        last_zero = target - 1 if apply else target
        if old > 0:
            for s_ in range(old + 1, last_zero + 1):
                exp_avg[r] *= beta1
                exp_avg_sq[r] *= beta2
                bc1, bc2 = 1 - beta1 ** s_, 1 - beta2 ** s_
                param[r] -= (lr / bc1) * exp_avg[r] / (exp_avg_sq[r].sqrt() / (bc2 ** 0.5) + eps)
        if apply:
            g = grad[r]
            if grad.dtype == torch.int64:
                g = g.to(torch.float32) * (1.0 / float(2 ** 48))
            exp_avg[r] = beta1 * exp_avg[r] + (1 - beta1) * g
            exp_avg_sq[r] = beta2 * exp_avg_sq[r] + (1 - beta2) * g * g
            bc1, bc2 = 1 - beta1 ** target, 1 - beta2 ** target
            param[r] -= (lr / bc1) * exp_avg[r] / (exp_avg_sq[r].sqrt() / (bc2 ** 0.5) + eps)
            grad[r] = 0
        row_stamp[r] = target


def install(monkeypatch):
    """Route savqa_b200.ops through these stand-ins for the duration of a test."""
    import savqa_b200.ops as real
    for name, fn in list(globals().items()):
        if callable(fn) and hasattr(real, name) and not name.startswith("_") and name not in ("install",):
            monkeypatch.setattr(real, name, fn)
