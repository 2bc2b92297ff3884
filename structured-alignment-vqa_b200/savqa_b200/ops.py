"""Typed Python wrappers over the C ABI (include/savqa_b200.h).  PyTorch is only the allocator / stream owner here:
every function validates shapes and dtypes, hands raw device pointers to libsavqa_b200.so and raises on error.
No function in this file computes anything with torch ops.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import AttnArgs, GemmEpilogue, GemmProblem, RowLnArgs, call, ptr

Tensor = torch.Tensor
BF16 = torch.bfloat16
F32 = torch.float32


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _rows2d(t: Tensor) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a tensor viewed as a row-major matrix over its last dim."""
    assert t.dim() >= 1 and t.stride(-1) == 1, "last dimension must be contiguous"
    if t.dim() == 1:
        return 1, t.shape[0], t.shape[0]
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0)
    assert t.is_contiguous(), "tensors of rank > 2 must be contiguous"
    return t.numel() // t.shape[-1], t.shape[-1], t.shape[-1]


def _check(t: Optional[Tensor], dtype, name: str) -> None:
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"savqa_b200: `{name}` must be a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"savqa_b200: `{name}` must be {dtype}, got {t.dtype}")


# ------------------------------------------------------------------------------------------------------
def build_masks(first_mask: Tensor, q_mask: Tensor, q_graph: Tensor, first_graph: Optional[Tensor], dec_mask_on: bool):
    """AttModel_x3.py:103-122 / :229-247 -> (graph_diag [B,T,T], graph [B,T,T], dec_mask [B,1,T]) fp32, bit exact."""
    is_float = first_mask.dtype == F32
    want = F32 if is_float else torch.int32
    ins = []
    for nm, t in (("first_mask", first_mask), ("q_mask", q_mask), ("q_graph", q_graph), ("first_graph", first_graph)):
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError(f"savqa_b200: `{nm}` must be a CUDA tensor (this path has no CPU fallback)")
            if t.dtype != want:
                t = t.to(want)
            t = t.contiguous()
        ins.append(t)
    fm, qm, qg, fg = ins
    B, V, Q = fm.shape[0], fm.shape[1], qm.shape[1]
    assert fm.shape == (B, V, V) and qm.shape == (B, Q, Q) and qg.shape == (B, Q, Q), "mask shapes"
    assert fg is None or fg.shape == (B, V, V)
    T = V + Q
    graph_diag = torch.empty(B, T, T, device=fm.device, dtype=F32)
    graph = torch.empty(B, T, T, device=fm.device, dtype=F32)
    dec_mask = torch.empty(B, 1, T, device=fm.device, dtype=F32)
    call("savqa_build_masks", ptr(fm), ptr(qm), ptr(qg), ptr(fg), int(is_float), B, V, Q, int(bool(dec_mask_on)),
         ptr(graph_diag), ptr(graph), ptr(dec_mask))
    # integer / bool inputs (what collate_fn emits: int32 0/1) give graphs that are exactly 0/1: hand the attention kernels their
    # bit-packed form as well (4 bytes per 32 keys).  Float inputs may carry weights (the reference multiplies by the graph
    # VALUES, modules.py:281-284): those stay on the fp32 graph path.
    if not is_float:
        for g in (graph_diag, graph):
            attach_graph_bits(g)
    return graph_diag, graph, dec_mask


def build_masks_compact(first_len: Tensor, q_len: Tensor, q_graph_bits: Tensor, first_graph_bits: Optional[Tensor], V: int, Q: int,
                        dec_mask_on: bool):
    """build_masks from the loader's compact hand-off (savqa_build_masks_compact): per-sample lengths (int32 [B]) instead of the
    prefix-block masks, bit-packed adjacency (int32 [B,Q,ceil(Q/32)], [B,V,ceil(V/32)] | None) instead of int32 planes.  Same three
    fp32 tensors as build_masks, bit for bit, with the graphs' bit-packed forms attached."""
    for nm, t in (("first_len", first_len), ("q_len", q_len), ("q_graph_bits", q_graph_bits), ("first_graph_bits", first_graph_bits)):
        _check(t, torch.int32, nm)
        assert t is None or t.is_contiguous()
    B = first_len.numel()
    assert q_len.numel() == B and q_graph_bits.shape == (B, Q, (Q + 31) // 32)
    assert first_graph_bits is None or first_graph_bits.shape == (B, V, (V + 31) // 32)
    T = V + Q
    dev = first_len.device
    graph_diag = torch.empty(B, T, T, device=dev, dtype=F32)
    graph = torch.empty(B, T, T, device=dev, dtype=F32)
    dec_mask = torch.empty(B, 1, T, device=dev, dtype=F32)
    wpr = (T + 31) // 32
    dbits = torch.empty(B, T, wpr, device=dev, dtype=torch.int32)
    gbits = torch.empty(B, T, wpr, device=dev, dtype=torch.int32)
    call("savqa_build_masks_compact", ptr(first_len), ptr(q_len), ptr(first_graph_bits), ptr(q_graph_bits), B, V, Q, int(bool(dec_mask_on)),
         ptr(graph_diag), ptr(graph), ptr(dec_mask), ptr(dbits), ptr(gbits))
    graph_diag._savqa_bits = (dbits, graph_diag._version, graph_diag.data_ptr())
    graph._savqa_bits = (gbits, graph._version, graph.data_ptr())
    return graph_diag, graph, dec_mask


def mil_nce_fwd(pn_h: Tensor, vis_h: Tensor, mask: Tensor, loc: Tensor, nodes: Tensor, B: int, V: int, M: int, topN: int, h: int):
    """MIL_NCE between its Linear layers (savqa_mil_nce_fwd).  Overwrites the object rows of `nodes` (bf16 [B*M, >= h]) with the
    refined object features; returns (raw fp32 [2, B*V*topN] for the backward, mil_nce_obj fp32 [1])."""
    _check(pn_h, BF16, "pn_h")
    _check(vis_h, BF16, "vis_h")
    _check(nodes, BF16, "nodes")
    _check(mask, torch.int32, "mask")
    _check(loc, torch.int64, "loc")
    assert pn_h.dim() == 2 and pn_h.shape[0] == 2 * B * V * topN and pn_h.shape[1] >= h and pn_h.stride(1) == 1
    assert vis_h.dim() == 2 and vis_h.shape[0] == B * V and vis_h.shape[1] >= h and vis_h.stride(1) == 1
    assert nodes.dim() == 2 and nodes.shape[0] == B * M and nodes.shape[1] >= h and nodes.stride(1) == 1
    assert mask.is_contiguous() and mask.numel() == B * V * topN and loc.is_contiguous() and loc.numel() == B * V
    raw = torch.empty(2, B * V * topN, device=pn_h.device, dtype=F32)
    term = torch.empty(B * V, device=pn_h.device, dtype=F32)
    obj = torch.empty(1, device=pn_h.device, dtype=F32)
    call("savqa_mil_nce_fwd", ptr(pn_h), pn_h.stride(0), ptr(vis_h), vis_h.stride(0), ptr(mask), ptr(loc), ptr(nodes), nodes.stride(0), B, V, M,
         topN, h, ptr(raw), ptr(term), ptr(obj))
    return raw, obj


def mil_nce_bwd(pn_h: Tensor, vis_h: Tensor, mask: Tensor, loc: Tensor, raw: Tensor, d_nodes: Optional[Tensor], d_obj: Optional[Tensor], B: int,
                V: int, M: int, topN: int, h: int):
    """Backward of mil_nce_fwd: returns the ReLU-gated pre-activation gradients (d_pn bf16 [2*B*V*topN, h], d_vis bf16 [B*V, h])."""
    _check(d_nodes, F32, "d_nodes")
    _check(d_obj, F32, "d_obj")
    assert d_nodes is None or (d_nodes.dim() == 2 and d_nodes.shape[0] == B * M and d_nodes.shape[1] >= h and d_nodes.stride(1) == 1)
    d_pn = torch.empty(2 * B * V * topN, h, device=pn_h.device, dtype=BF16)
    d_vis = torch.empty(B * V, h, device=pn_h.device, dtype=BF16)
    call("savqa_mil_nce_bwd", ptr(pn_h), pn_h.stride(0), ptr(vis_h), vis_h.stride(0), ptr(mask), ptr(loc), ptr(raw), ptr(d_nodes),
         d_nodes.stride(0) if d_nodes is not None else 0, ptr(d_obj), B, V, M, topN, h, ptr(d_pn), h, ptr(d_vis), h)
    return d_pn, d_vis


def pack_graph_bits(graph: Tensor) -> Tensor:
    """int32 [N, Tq, ceil(Tk / 32)]: bit j of word w of a row = (graph[row, 32 w + j] != 0).  For 0/1 graphs only."""
    _check(graph, F32, "graph")
    assert graph.dim() == 3 and graph.is_contiguous()
    N, Tq, Tk = graph.shape
    wpr = (Tk + 31) // 32
    bits = torch.empty(N, Tq, wpr, device=graph.device, dtype=torch.int32)
    call("savqa_pack_graph_bits", ptr(graph), N * Tq, Tk, ptr(bits), wpr)
    return bits


def attach_graph_bits(graph: Tensor) -> Tensor:
    """Marks a 0/1 graph tensor with its bit-packed form; graph_bits_of() returns it while the tensor is unmodified."""
    graph._savqa_bits = (pack_graph_bits(graph), graph._version, graph.data_ptr())
    return graph


def graph_bits_of(graph: Optional[Tensor]) -> Optional[Tensor]:
    tag = getattr(graph, "_savqa_bits", None) if graph is not None else None
    if tag is not None and tag[1] == graph._version and tag[2] == graph.data_ptr():
        return tag[0]
    return None


def gather_rows(table: Tensor, idx: Tensor, scale: float = 1.0, want_f32: bool = True, want_bf16: bool = False):
    """out[r] = table[idx[r]] * scale.  Returns (fp32 [n, width] | None, bf16 [n, pad8(width)] | None)."""
    _check(table, F32, "table")
    _check(idx, torch.int64, "idx")
    assert table.dim() == 2 and table.is_contiguous()
    idx = idx.contiguous()
    n, width = idx.numel(), table.shape[1]
    out32 = torch.empty(n, width, device=table.device, dtype=F32) if want_f32 else None
    ldb = pad8(width)
    out16 = torch.empty(n, ldb, device=table.device, dtype=BF16) if want_bf16 else None
    call("savqa_gather_rows", ptr(table), table.shape[0], width, ptr(idx), n, float(scale), ptr(out32), width, ptr(out16), ldb, ldb)
    return out32, out16


Q48 = float(2 ** 48)  # scale of the fixed-point row-gradient accumulator (savqa_scatter_add_rows_q48)


def scatter_add_rows(dtable: Tensor, idx: Tensor, dout: Tensor, scale: float = 1.0, skip_row: int = -1) -> None:
    """dtable[idx[r]] += dout[r] * scale.  A float32 dtable takes float atomics (savqa_scatter_add_rows); an int64 dtable is the Q15.48
    fixed-point accumulator whose result does not depend on the order of the atomics (savqa_scatter_add_rows_q48) -- the form the
    trainer uses, so that data-parallel replicas stay bit-identical."""
    assert dtable.dtype in (F32, torch.int64), f"dtable: float32 or int64 (Q15.48) expected, got {dtable.dtype}"
    _check(idx, torch.int64, "idx")
    _check(dout, F32, "dout")
    idx = idx.contiguous()
    rows, width, ld = _rows2d(dout)
    assert rows == idx.numel() and width == dtable.shape[1] and dtable.is_contiguous()
    fn = "savqa_scatter_add_rows" if dtable.dtype == F32 else "savqa_scatter_add_rows_q48"
    call(fn, ptr(dtable), dtable.shape[0], width, ptr(idx), rows, ptr(dout), ld, float(scale), int(skip_row))


def cast_bf16(src: Tensor, out: Optional[Tensor] = None, pad_to: Optional[int] = None) -> Tensor:
    """fp32 [rows, cols] -> bf16 [rows, pad8(cols)] (pad columns zero)."""
    _check(src, F32, "src")
    rows, cols, ld = _rows2d(src)
    if out is None:
        pad_to = pad8(cols) if pad_to is None else pad_to
        out = torch.empty(rows, pad_to, device=src.device, dtype=BF16)
    else:
        _check(out, BF16, "out")
        pad_to = cols if pad_to is None else pad_to
    orows, ocols, old = _rows2d(out)
    assert orows == rows and ocols >= pad_to
    call("savqa_cast_bf16", ptr(src), ld, ptr(out), old, rows, cols, pad_to)
    return out


def cast_transpose_bf16(src: Tensor, out: Tensor, pad_to: Optional[int] = None) -> Tensor:
    """out[c, r] = src[r, c];  out is a (possibly strided) bf16 matrix [cols, >= rows]."""
    _check(src, F32, "src")
    _check(out, BF16, "out")
    rows, cols, ld = _rows2d(src)
    orows, ocols, old = _rows2d(out)
    pad_to = rows if pad_to is None else pad_to
    assert orows == cols and ocols >= pad_to >= rows
    call("savqa_cast_transpose_bf16", ptr(src), ld, ptr(out), old, rows, cols, pad_to)
    return out


def row_nonzero(x: Tensor, want_bf16: bool = True):
    """on[r] = (sum_c x[r,c] != 0) -- the key/query masks of modules.py:257,289 -- plus a bf16 copy of x."""
    _check(x, F32, "x")
    rows, cols, ld = _rows2d(x)
    on = torch.empty(rows, device=x.device, dtype=F32)
    xb = torch.empty(rows, pad8(cols), device=x.device, dtype=BF16) if want_bf16 else None
    if xb is not None and pad8(cols) != cols:
        xb.zero_()
    call("savqa_row_nonzero", ptr(x), ld, rows, cols, ptr(on), ptr(xb), pad8(cols))
    return on, xb


def relu_gate_bf16(dy: Tensor, act: Tensor, group_rows: int = 0, group_stride: int = 0) -> Tensor:
    """bf16 (act > 0 ? dy : 0); dy fp32 or bf16, act bf16 -- the staged A operand of dgrad / wgrad behind a ReLU.
    group_rows > 0: dy is the [*, ld] matrix that holds the wanted rows in groups of `group_rows` every `group_stride` rows."""
    _check(act, BF16, "act")
    if dy.dtype not in (F32, BF16) or not dy.is_cuda:
        raise TypeError("savqa_b200: `dy` must be a CUDA fp32 or bf16 tensor")
    arows, acols, ald = _rows2d(act)
    if group_rows > 0:
        assert dy.dim() == 2 and dy.stride(1) == 1
        rows, cols, ld = arows, dy.shape[1], dy.stride(0)
        assert rows % group_rows == 0 and dy.shape[0] >= (rows // group_rows - 1) * group_stride + group_rows
    else:
        rows, cols, ld = _rows2d(dy)
    assert arows == rows and acols >= cols
    out = torch.empty(rows, pad8(cols), device=dy.device, dtype=BF16)
    if pad8(cols) != cols:
        out.zero_()
    call("savqa_relu_gate_bf16", ptr(dy), int(dy.dtype == F32), ld, ptr(act), ald, ptr(out), out.stride(0), rows, cols, int(group_rows),
         int(group_stride))
    return out


def regroup_cols(x: Tensor, groups: int, w_in: int, w_out: int, out: Optional[Tensor] = None) -> Tensor:
    """[rows, groups * w_in] -> [rows, groups * w_out]: every group's first min(w_in, w_out) columns copied, the rest zero
    (savqa_regroup_cols).  x / out may be column slices of wider matrices (row pitch = stride(0)); bf16 or fp32."""
    if x.dtype not in (F32, BF16) or not x.is_cuda:
        raise TypeError("savqa_b200: `x` must be a CUDA fp32 or bf16 tensor")
    assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] >= groups * w_in
    if out is None:
        out = torch.empty(x.shape[0], groups * w_out, device=x.device, dtype=x.dtype)
    assert out.dtype == x.dtype and out.dim() == 2 and out.stride(1) == 1 and out.shape[0] == x.shape[0] and out.shape[1] >= groups * w_out
    call("savqa_regroup_cols", ptr(x), x.stride(0), ptr(out), out.stride(0), x.shape[0], groups, w_in, w_out, x.element_size())
    return out


def fill_zero(t: Tensor, max_blocks: int = 0) -> None:
    """t.zero_() by a bounded number of persistent blocks (savqa_fill_zero): a background fill under other streams' kernels."""
    assert t.is_cuda and t.is_contiguous()
    call("savqa_fill_zero", ptr(t), t.numel() * t.element_size(), int(max_blocks))


def colsum_bf16(x: Tensor, out: Tensor) -> None:
    """out[c] += sum_r x[r, c]"""
    _check(x, BF16, "x")
    _check(out, F32, "out")
    rows, cols, ld = _rows2d(x)
    assert out.numel() == cols and out.is_contiguous()
    call("savqa_colsum_bf16", ptr(x), ld, rows, cols, ptr(out))


def layernorm_fwd(x: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float, save_pre: bool, want_bf16: bool,
                  want_on: bool, stats: Optional[Tensor] = None):
    """y = LN(x + res) (modules.py:62-65).  Returns (y, pre | None, y_bf16 | None, on | None); `stats` (fp32 [rows, 2], optional)
    receives {mean, sigma} of every row."""
    for nm, t in (("x", x), ("res", res), ("gamma", gamma), ("beta", beta)):
        _check(t, F32, nm)
    assert x.is_contiguous() and (res is None or (res.is_contiguous() and res.shape == x.shape))
    C_ = x.shape[-1]
    rows = x.numel() // C_
    y = torch.empty_like(x)
    pre = torch.empty_like(x) if save_pre else None
    yb = torch.empty(x.shape, device=x.device, dtype=BF16) if want_bf16 else None
    on = torch.empty(rows, device=x.device, dtype=F32) if want_on else None
    _check(stats, F32, "stats")
    assert stats is None or (stats.is_contiguous() and stats.numel() >= 2 * rows)
    call("savqa_residual_layernorm_fwd", ptr(x), ptr(res), ptr(gamma), ptr(beta), float(eps), rows, C_, ptr(pre), ptr(y), ptr(yb), ptr(on),
         ptr(stats))
    return y, pre, yb, on


def layernorm_bwd(dy: Tensor, pre: Tensor, gamma: Tensor, eps: float, dgamma: Optional[Tensor], dbeta: Optional[Tensor],
                  dres_in: Optional[Tensor] = None, want_bf16: bool = False, dxsum: Optional[Tensor] = None):
    """Returns (dx fp32, dx_bf16 | None); dgamma / dbeta / dxsum (column sums of dx) are accumulated in place."""
    for nm, t in (("dy", dy), ("pre", pre), ("gamma", gamma), ("dgamma", dgamma), ("dbeta", dbeta), ("dres_in", dres_in), ("dxsum", dxsum)):
        _check(t, F32, nm)
    assert dy.is_contiguous() and pre.is_contiguous() and dy.shape == pre.shape
    C_ = pre.shape[-1]
    rows = pre.numel() // C_
    dx = torch.empty_like(pre)
    dxb = torch.empty(pre.shape, device=pre.device, dtype=BF16) if want_bf16 else None
    for t in (dgamma, dbeta, dxsum):
        assert t is None or (t.is_contiguous() and t.numel() == C_)
    call("savqa_layernorm_bwd", ptr(dy), ptr(pre), ptr(gamma), float(eps), rows, C_, ptr(dres_in), ptr(dx), ptr(dxb), ptr(dgamma), ptr(dbeta),
         ptr(dxsum))
    return dx, dxb


# ------------------------------------------------------------------------------------------------------


#: experiment (SAVQA_BRANCH_SMS): cuda stream handle -> SMs the persistent GEMMs launched on that stream may take.  The two branch
#: models run on two streams; with a static split of the SMs between them (in proportion to their work) a GEMM of one branch computes
#: while the other branch's GEMM pays its fixed launch / pipeline-fill / tail cost, instead of the two taking turns on all 148 SMs.
_STREAM_SM_LIMIT = {}


def set_stream_sm_limit(stream, sms: int) -> None:
    if sms and sms > 0:
        _STREAM_SM_LIMIT[stream.cuda_stream] = int(sms)
    else:
        _STREAM_SM_LIMIT.pop(stream.cuda_stream, None)


def clear_stream_sm_limits() -> None:
    """Forgets every per-stream SM budget (the branch models set theirs again at their next forward): for code that times a
    GEMM alone on a stream a model has run on."""
    _STREAM_SM_LIMIT.clear()


@contextlib.contextmanager
def _stream_limit():
    lim = _STREAM_SM_LIMIT.get(torch.cuda.current_stream().cuda_stream) if _STREAM_SM_LIMIT else None
    if not lim:
        yield
        return
    lib = _lib.load()
    old = lib.savqa_set_gemm_sm_limit(int(lim))
    if old and old < lim:
        lib.savqa_set_gemm_sm_limit(old)
    try:
        yield
    finally:
        lib.savqa_set_gemm_sm_limit(old)


def gemm(a: Tensor, b: Tensor, M: int, N: int, K: int, *, a_mn: bool = False, b_mn: bool = False, bias: Optional[Tensor] = None,
         res: Optional[Tensor] = None, rowtab: Optional[Tensor] = None, rowtab_period: int = 0, gate: Optional[Tensor] = None,
         relu: bool = False, alpha: float = 1.0, out_f32: Optional[Tensor] = None, out_bf16: Optional[Tensor] = None,
         accumulate: int = 0, split_k: int = 1, colsum: Optional[Tensor] = None) -> None:
    """acc[m,n] = sum_k A[m,k] B[n,k] with the fused epilogue of savqa_gemm_bf16 (tcgen05 / TMEM / TMA kernel).

    K-major operands are [rows, >=K] matrices; MN-major operands ([K, >=rows]) are used by wgrad."""
    _check(a, BF16, "A")
    _check(b, BF16, "B")
    for nm, t in (("bias", bias), ("res", res), ("rowtab", rowtab), ("out_f32", out_f32), ("colsum", colsum)):
        _check(t, F32, nm)
    _check(gate, BF16, "gate")
    _check(out_bf16, BF16, "out_bf16")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_mn:
        assert a.shape[0] >= K and a.shape[1] >= M, (a.shape, M, K)
    else:
        assert a.shape[0] >= M and a.shape[1] >= K, (a.shape, M, K)
    if b_mn:
        assert b.shape[0] >= K and b.shape[1] >= N, (b.shape, N, K)
    else:
        assert b.shape[0] >= N and b.shape[1] >= K, (b.shape, N, K)
    e = GemmEpilogue()
    _fill_epilogue(e, M, N, bias, res, rowtab, rowtab_period, gate, relu, alpha, out_f32, out_bf16, accumulate, split_k, colsum)
    with _stream_limit():
        call("savqa_gemm_bf16", ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn), M, N, K, C.byref(e), int(split_k))


def _fill_epilogue(e, M, N, bias, res, rowtab, rowtab_period, gate, relu, alpha, out_f32, out_bf16, accumulate, split_k, colsum) -> None:
    e.alpha, e.relu, e.accumulate, e.rowtab_period = float(alpha), int(relu), int(accumulate), int(rowtab_period)
    e.bias = ptr(bias)
    if bias is not None:
        assert bias.numel() >= N and bias.is_contiguous()
    if res is not None:
        assert res.dim() == 2 and res.stride(1) == 1 and res.shape[0] >= M and res.shape[1] >= N
        e.res, e.ld_res = ptr(res), res.stride(0)
    if rowtab is not None:
        assert rowtab.dim() == 2 and rowtab.stride(1) == 1 and rowtab.shape[0] >= rowtab_period > 0 and rowtab.shape[1] >= N
        e.rowtab, e.ld_rowtab = ptr(rowtab), rowtab.stride(0)
    if gate is not None:
        assert gate.dim() == 2 and gate.stride(1) == 1 and gate.shape[0] >= M and gate.shape[1] >= N
        e.gate_bf16, e.ld_gate = ptr(gate), gate.stride(0)
    if out_f32 is not None:
        assert out_f32.dim() == 2 and out_f32.stride(1) == 1 and out_f32.shape[0] >= M and out_f32.shape[1] >= N
        e.out_f32, e.ld_out_f32 = ptr(out_f32), out_f32.stride(0)
    if out_bf16 is not None:
        assert out_bf16.dim() == 2 and out_bf16.stride(1) == 1 and out_bf16.shape[0] >= M and out_bf16.shape[1] >= N
        e.out_bf16, e.ld_out_bf16 = ptr(out_bf16), out_bf16.stride(0)
    if colsum is not None:
        assert colsum.is_contiguous() and colsum.numel() >= N and split_k == 1
        e.colsum = ptr(colsum)


def rowln_fits(M: int, N: int, K: int) -> bool:
    """Shapes savqa_gemm_rowln's row-wise (LayerNorm) modes take: one cluster of N / 64 <= 8 CTAs per 128-row block."""
    return N in (64, 128, 256, 512) and K % 8 == 0 and M >= 1


def gemm_rowln(a: Tensor, b: Tensor, M: int, N: int, K: int, mode: int, *, b_mn: bool = False, bias: Optional[Tensor] = None,
               relu: bool = False, rowscale: Optional[Tensor] = None, res: Optional[Tensor] = None, gate: Optional[Tensor] = None,
               gamma: Optional[Tensor] = None, beta: Optional[Tensor] = None, eps: float = 1e-8, act_bf16: Optional[Tensor] = None,
               pre: Optional[Tensor] = None, y: Optional[Tensor] = None, y_bf16: Optional[Tensor] = None, on: Optional[Tensor] = None,
               stats: Optional[Tensor] = None, dxg_bf16: Optional[Tensor] = None, dgamma: Optional[Tensor] = None,
               dbeta: Optional[Tensor] = None, dxsum: Optional[Tensor] = None) -> None:
    """Cluster GEMM with a row-wise epilogue (savqa_gemm_rowln): mode 0 plain, 1 Linear -> (* rowscale) -> + res -> LayerNorm,
    2 dgrad -> + res -> LayerNorm backward.  All matrices are 2-D views with a contiguous last dimension."""
    _check(a, BF16, "A")
    _check(b, BF16, "B")
    for nm, t in (("bias", bias), ("rowscale", rowscale), ("res", res), ("gamma", gamma), ("beta", beta), ("pre", pre), ("y", y), ("on", on),
                  ("stats", stats), ("dgamma", dgamma), ("dbeta", dbeta), ("dxsum", dxsum)):
        _check(t, F32, nm)
    for nm, t in (("gate", gate), ("act_bf16", act_bf16), ("y_bf16", y_bf16), ("dxg_bf16", dxg_bf16)):
        _check(t, BF16, nm)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1 and a.shape[0] >= M and a.shape[1] >= K
    assert (b.shape[0] >= K and b.shape[1] >= N) if b_mn else (b.shape[0] >= N and b.shape[1] >= K)
    g = RowLnArgs()
    g.A, g.lda, g.B, g.ldb, g.b_mn_major = ptr(a), a.stride(0), ptr(b), b.stride(0), int(b_mn)
    g.M, g.N, g.K, g.mode, g.relu = M, N, K, int(mode), int(relu)
    g.bias, g.rowscale, g.gamma, g.beta, g.eps = ptr(bias), ptr(rowscale), ptr(gamma), ptr(beta), float(eps)
    g.on, g.stats, g.dgamma, g.dbeta, g.dxsum = ptr(on), ptr(stats), ptr(dgamma), ptr(dbeta), ptr(dxsum)
    for t in (bias, gamma, beta, dgamma, dbeta, dxsum):
        assert t is None or (t.is_contiguous() and t.numel() >= N)
    for t in (rowscale, on):
        assert t is None or (t.is_contiguous() and t.numel() >= M)
    assert stats is None or (stats.is_contiguous() and stats.numel() >= 2 * M)

    def mat(t, nm):
        assert t.dim() == 2 and t.stride(1) == 1 and t.shape[0] >= M and t.shape[1] >= N, nm
        return ptr(t), t.stride(0)
    if res is not None:
        g.res, g.ld_res = mat(res, "res")
    if gate is not None:
        g.gate_bf16, g.ld_gate = mat(gate, "gate")
    if act_bf16 is not None:
        g.act_bf16, g.ld_act = mat(act_bf16, "act_bf16")
    if pre is not None:
        g.pre, g.ld_pre = mat(pre, "pre")
    if y is not None:
        g.y, g.ld_y = mat(y, "y")
    if y_bf16 is not None:
        g.y_bf16, g.ld_yb = mat(y_bf16, "y_bf16")
    if dxg_bf16 is not None:
        g.dxg_bf16, g.ld_dxg = mat(dxg_bf16, "dxg_bf16")
    call("savqa_gemm_rowln", C.byref(g))


@contextlib.contextmanager
def gemm_sm_limit(sms: int):
    """The GEMMs launched inside (by this thread) leave 148 - sms SMs to kernels of other streams: savqa_set_gemm_sm_limit."""
    if not sms:
        yield
        return
    lib = _lib.load()
    old = lib.savqa_set_gemm_sm_limit(int(sms))
    try:
        yield
    finally:
        lib.savqa_set_gemm_sm_limit(old)


def gemm_grouped(problems, N: int, *, a_mn: bool = False, b_mn: bool = False, split_k: int = 1) -> None:
    """Several independent GEMMs with the same N / operand majors / kind of output in ONE launch (savqa_gemm_bf16_grouped): the
    same layer of the two branch models.  `problems`: dicts with a, b, M, K and the epilogue keywords of gemm()."""
    if len(problems) == 1 or len(problems) > 2:
        for p in problems:
            kw = {k: v for k, v in p.items() if k not in ("a", "b", "M", "K")}
            gemm(p["a"], p["b"], p["M"], N, p["K"], a_mn=a_mn, b_mn=b_mn, split_k=split_k, **kw)
        return
    arr = (GemmProblem * len(problems))()
    for i, p in enumerate(problems):
        a, b, M, K = p["a"], p["b"], p["M"], p["K"]
        _check(a, BF16, "A")
        _check(b, BF16, "B")
        assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
        for nm in ("bias", "res", "rowtab", "out_f32", "colsum"):
            _check(p.get(nm), F32, nm)
        _check(p.get("gate"), BF16, "gate")
        _check(p.get("out_bf16"), BF16, "out_bf16")
        arr[i].A, arr[i].lda, arr[i].B, arr[i].ldb, arr[i].M, arr[i].K = ptr(a), a.stride(0), ptr(b), b.stride(0), M, K
        _fill_epilogue(arr[i].epilogue, M, N, p.get("bias"), p.get("res"), p.get("rowtab"), p.get("rowtab_period", 0), p.get("gate"),
                       p.get("relu", False), p.get("alpha", 1.0), p.get("out_f32"), p.get("out_bf16"), p.get("accumulate", 0), split_k,
                       p.get("colsum"))
    call("savqa_gemm_bf16_grouped", arr, len(problems), int(a_mn), int(b_mn), N, int(split_k))


def wgrad_grouped(items) -> None:
    """[(dy, x, n_out, k_in, out)], same n_out / k_in: out += dy^T x for each, in one launch."""
    n_out, k_in = items[0][2], items[0][3]
    tiles = ((n_out + 127) // 128) * ((k_in + 127) // 128)
    sk = min(split_k_for(tiles, (it[0].shape[0] + 63) // 64) for it in items)
    gemm_grouped([dict(a=dy, b=x, M=n_out, K=dy.shape[0], out_f32=out, accumulate=2) for dy, x, _, _, out in items], k_in,
                 a_mn=True, b_mn=True, split_k=sk)


def split_k_for(tiles: int, k_blocks: int) -> int:
    """Enough K-splits to put ~2 work units on every SM, at least 4 k-blocks each."""
    if tiles >= 2 * 148:
        return 1
    want = (2 * 148 + tiles - 1) // tiles
    return max(1, min(want, k_blocks // 4 if k_blocks >= 8 else 1))


def wgrad(dy: Tensor, x: Tensor, n_out: int, k_in: int, out: Tensor) -> None:
    """out[n_out, k_in] += dy[M, n_out]^T x[M, k_in]  (both operands read MN-major straight from the activations)."""
    M = dy.shape[0]
    assert x.shape[0] == M
    tiles = ((n_out + 127) // 128) * ((k_in + 127) // 128)
    gemm(dy, x, n_out, k_in, M, a_mn=True, b_mn=True, out_f32=out, accumulate=2, split_k=split_k_for(tiles, (M + 63) // 64))


# ------------------------------------------------------------------------------------------------------
def _set_graph_bits(a: AttnArgs, graph_bits: Optional[Tensor], N: int, Tq: int, Tk: int) -> None:
    if graph_bits is None:
        return
    _check(graph_bits, torch.int32, "graph_bits")
    wpr = (Tk + 31) // 32
    assert graph_bits.is_contiguous() and graph_bits.shape[0] == N and graph_bits.shape[1] in (Tq, 1) and graph_bits.shape[2] == wpr
    a.graph_bits, a.bits_n_stride = ptr(graph_bits), graph_bits.shape[1] * wpr
    a.bits_q_stride = wpr if graph_bits.shape[1] == Tq else 0


def _adjacent(a: Tensor, b: Tensor, c: Tensor, width: int) -> bool:
    """a | b | c are consecutive `width`-column blocks of one row-major matrix (the fused [M, 3C] projection layout)."""
    es = a.element_size()
    return (a.shape[0] == b.shape[0] == c.shape[0] and a.stride(0) == b.stride(0) == c.stride(0) and a.stride(0) >= 3 * width
            and b.data_ptr() == a.data_ptr() + width * es and c.data_ptr() == b.data_ptr() + width * es)


def _pad_heads_qkv(q: Tensor, k: Tensor, v: Tensor, H: int):
    """32-channel heads -> zero-padded 64-wide heads; one launch when q | k | v are the column blocks of one fused projection."""
    if _adjacent(q, k, v, H * 32):
        fused = torch.as_strided(q, (q.shape[0], 3 * H * 32), (q.stride(0), 1))
        out = regroup_cols(fused, 3 * H, 32, 64)
        return out[:, :H * 64], out[:, H * 64:2 * H * 64], out[:, 2 * H * 64:]
    return tuple(regroup_cols(t, H, 32, 64) for t in (q, k, v))


def graph_attention_fwd(q: Tensor, k: Tensor, v: Tensor, graph: Optional[Tensor], key_on: Tensor, query_on: Tensor, N: int, H: int,
                        Tq: int, Tk: int, d: int, causal: bool, renorm: int, want_att: bool, engine: int,
                        graph_bits: Optional[Tensor] = None, stats: Optional[Tensor] = None, scale_d: int = 0):
    """Attention core of modules.py:246-301.  q/k/v: bf16 2-D views [N*T, >= H*d].  Returns (out fp32 [N*Tq, H*d], att | None)."""
    for nm, t in (("q", q), ("k", k), ("v", v)):
        _check(t, BF16, nm)
        assert t.dim() == 2 and t.stride(1) == 1
    if engine == 0 and d == 32:
        # 32-channel heads (16 heads x 512) on the tcgen05 engine: 64-wide tiles whose upper halves are zero; the score scale stays
        # 1/sqrt(32) (AttnArgs.scale_d).  The zero channels change neither Q K^T nor the first 32 columns of P V.
        q64, k64, v64 = _pad_heads_qkv(q, k, v, H)
        out64, att = graph_attention_fwd(q64, k64, v64, graph, key_on, query_on, N, H, Tq, Tk, 64, causal, renorm, want_att, 0,
                                         graph_bits=graph_bits, stats=stats, scale_d=32)
        return regroup_cols(out64, H, 64, 32), att
    _check(graph, F32, "graph")
    a = AttnArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0)
    if graph is not None:
        assert graph.dim() == 3 and graph.shape[0] == N and graph.shape[2] == Tk and graph.shape[1] in (Tq, 1) and graph.is_contiguous()
        a.graph, a.graph_n_stride = ptr(graph), graph.shape[1] * Tk
        a.graph_q_stride = Tk if graph.shape[1] == Tq else 0
    a.key_on, a.query_on = ptr(key_on), ptr(query_on)
    a.N, a.H, a.Tq, a.Tk, a.d = N, H, Tq, Tk, d
    a.causal, a.renorm, a.engine, a.scale_d = int(causal), int(renorm), int(engine), int(scale_d)
    if graph is not None:
        _set_graph_bits(a, graph_bits, N, Tq, Tk)
    out = torch.empty(N * Tq, H * d, device=q.device, dtype=F32)
    att = torch.empty(H * N, Tq, Tk, device=q.device, dtype=F32) if want_att else None
    a.out, a.ldo, a.att = ptr(out), H * d, ptr(att)
    if stats is not None:  # engine 0 only: {m, +-1/Z, scale, beta} per (head, sample, query) row, for the backward kernel
        _check(stats, F32, "stats")
        assert engine == 0 and stats.is_contiguous() and stats.numel() == H * N * Tq * 4
        a.stats = ptr(stats)
    call("savqa_graph_attn_fwd", C.byref(a))
    return out, att


def _bwd_one_cta_fits(d: int, Tq: int, Tk: int) -> bool:
    """One CTA of csrc/attn_bwd_tcgen05.cu holds the whole (sample, head) problem: Q, dO, K, V and the two bf16 [128, Tk] tiles in
    shared memory, S / dW / dQ / dK / dV in 512 TMEM columns."""
    if d not in (64, 128) or Tq > 128 or Tk > 256:
        return False
    kt = (Tk + 127) // 128
    tk16 = (Tk + 15) // 16 * 16
    dw_off = (tk16 + 31) // 32 * 32
    if max(dw_off + tk16, (1 + 2 * kt) * d) > 512:
        return False
    dch = d // 64
    smem = 1024 + 2 * dch * 16384 + 2 * dch * tk16 * 128 + 2 * ((tk16 + 63) // 64) * 16384 + Tk * 4 + 16 + 128 * ((Tk + 31) // 32) * 4
    return smem + 64 <= 227 * 1024


def tc_attention_bwd_fits(d: int, Tq: int, Tk: int) -> bool:
    """Shapes the tcgen05 attention backward takes: directly when one CTA holds the (sample, head) problem, otherwise as (query tile,
    key tile) pairs of <= 128 x 128 on the forward's row statistics (graph_attention_bwd); 32-channel heads as zero-padded 64-wide
    tiles."""
    if d == 32:
        d = 64
    if d not in (64, 128):
        return False
    return _bwd_one_cta_fits(d, Tq, Tk) or (Tq <= 512 and Tk <= 512 and _bwd_one_cta_fits(d, min(Tq, 128), min(Tk, 128)))


def graph_attention_bwd(q, k, v, graph, key_on, query_on, N, H, Tq, Tk, d, causal, renorm, dout: Tensor, dq: Tensor, dk: Tensor,
                        dv: Tensor, engine: Optional[int] = None, dbq: Optional[Tensor] = None, dbk: Optional[Tensor] = None,
                        dbv: Optional[Tensor] = None, graph_bits: Optional[Tensor] = None, stats: Optional[Tensor] = None,
                        fwd_out: Optional[Tensor] = None, scale_d: int = 0) -> None:
    """Gradient of the attention core; dq/dk/dv are bf16 2-D views and come back ReLU-gated by q/k/v > 0.
    dbq/dbk/dbv (fp32 [H*d], optional) accumulate the column sums of dq/dk/dv: the projections' bias gradients.
    engine None: tcgen05 kernel when the shape fits, CUDA-core kernels otherwise (Tq == 1: the one-warp row kernel)."""
    if engine is None:
        strides_ok = all(t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0 for t in (q, k, v, dq, dk, dv))
        engine = 0 if (tc_attention_bwd_fits(d, Tq, Tk) and strides_ok and Tq > 1) else 1
    if engine == 0 and d == 32:
        # zero-padded 64-wide heads on the tcgen05 engine (see graph_attention_fwd); the gradients of the padding are dropped again
        C32 = H * 32
        q64, k64, v64 = _pad_heads_qkv(q, k, v, H)
        dout64 = regroup_cols(dout, H, 32, 64)
        fo64 = regroup_cols(fwd_out, H, 32, 64) if fwd_out is not None else None
        fused = _adjacent(dq, dk, dv, C32)
        if fused:  # dq | dk | dv are the three column blocks of one [M, 3C] gradient: one padded buffer, one launch back
            all64 = torch.empty(dq.shape[0], 3 * H * 64, device=q.device, dtype=BF16)
            d64 = [all64[:, i * H * 64:(i + 1) * H * 64] for i in range(3)]
        else:
            d64 = [torch.empty(t.shape[0], H * 64, device=q.device, dtype=BF16) for t in (dq, dk, dv)]
        graph_attention_bwd(q64, k64, v64, graph, key_on, query_on, N, H, Tq, Tk, 64, causal, renorm, dout64, d64[0], d64[1], d64[2], engine=0,
                            graph_bits=graph_bits, stats=stats, fwd_out=fo64, scale_d=32)
        if fused:
            regroup_cols(all64, 3 * H, 64, 32, out=torch.as_strided(dq, (dq.shape[0], 3 * C32), (dq.stride(0), 1)))
        for src, dst, db in zip(d64, (dq, dk, dv), (dbq, dbk, dbv)):
            if not fused:
                regroup_cols(src, H, 64, 32, out=dst)
            if db is not None:
                colsum_bf16(dst[:, :C32], db)
        return
    if engine == 0 and (Tq > 128 or not _bwd_one_cta_fits(d, Tq, Tk)):
        # The tcgen05 backward holds one (<= 128 query) x (all keys) problem per CTA.  Larger problems are TILED here: with the forward's
        # row statistics {m, 1/Z, scale, beta} and t = <dO, O> every (query tile, key tile) pair is independent -- W' and dS of a tile
        # need nothing from the other key tiles -- so each pair is one launch on contiguous copies of its rows; dQ sums over the key
        # tiles, dK / dV over the query tiles (the ReLU gates are per element, so the tiles' gated contributions are additive).
        assert stats is not None and fwd_out is not None and not causal, "tiled tcgen05 backward needs the forward statistics"
        Cq = H * d
        qt = [(t0, min(Tq, t0 + 128)) for t0 in range(0, Tq, 128)]
        if _bwd_one_cta_fits(d, min(Tq, 128), Tk):
            kstep = Tk
        else:  # balanced key tiles that start on 32-key (bit-word) boundaries
            nk = (Tk + 127) // 128
            kstep = ((Tk + nk - 1) // nk + 31) // 32 * 32
        kt = [(s0, min(Tk, s0 + kstep)) for s0 in range(0, Tk, kstep)]
        rows = lambda t, T, t0, t1, w: t.unflatten(0, (N, T))[:, t0:t1, :w].reshape(N * (t1 - t0), w).contiguous()  # noqa: E731
        dq3, dk3, dv3 = dq.unflatten(0, (N, Tq)), dk.unflatten(0, (N, Tk)), dv.unflatten(0, (N, Tk))
        st4 = stats.reshape(H, N, Tq, 4)
        for qi, (t0, t1) in enumerate(qt):
            n = t1 - t0
            q_t, do_t, fo_t = rows(q, Tq, t0, t1, Cq), rows(dout, Tq, t0, t1, Cq), rows(fwd_out, Tq, t0, t1, Cq)
            st_t = st4[:, :, t0:t1].contiguous().reshape(-1, 4)
            qon_t = query_on.reshape(N, Tq)[:, t0:t1].contiguous()
            dq_acc = None
            for ki, (s0, s1) in enumerate(kt):
                m = s1 - s0
                whole = (m == Tk)
                k_t, v_t = (k, v) if whole else (rows(k, Tk, s0, s1, Cq), rows(v, Tk, s0, s1, Cq))
                kon_t = key_on if whole else key_on.reshape(N, Tk)[:, s0:s1].contiguous()
                g_t, gb_t = graph, graph_bits
                if graph is not None:
                    gq = slice(t0, t1) if graph.shape[1] == Tq else slice(None)
                    g_t = graph[:, gq, s0:s1].contiguous()
                    gb_t = graph_bits[:, gq, s0 // 32:(s1 + 31) // 32].contiguous() if graph_bits is not None else None
                dq_t = torch.empty(N * n, Cq, device=q.device, dtype=BF16)
                dk_t, dv_t = torch.empty(N * m, Cq, device=q.device, dtype=BF16), torch.empty(N * m, Cq, device=q.device, dtype=BF16)
                graph_attention_bwd(q_t, k_t, v_t, g_t, kon_t, qon_t, N, H, n, m, d, False, renorm, do_t, dq_t, dk_t, dv_t, engine=0,
                                    graph_bits=gb_t, stats=st_t, fwd_out=fo_t, scale_d=scale_d or d)
                dq_acc = dq_t if dq_acc is None else dq_acc.add_(dq_t)
                tgt_k, tgt_v = dk3[:, s0:s1, :Cq], dv3[:, s0:s1, :Cq]
                if qi == 0:
                    tgt_k.copy_(dk_t.reshape(N, m, Cq))
                    tgt_v.copy_(dv_t.reshape(N, m, Cq))
                else:
                    tgt_k.add_(dk_t.reshape(N, m, Cq))
                    tgt_v.add_(dv_t.reshape(N, m, Cq))
            dq3[:, t0:t1, :Cq].copy_(dq_acc.reshape(N, n, Cq))
        for g_, db in ((dq, dbq), (dk, dbk), (dv, dbv)):
            if db is not None:
                colsum_bf16(g_[:, :Cq], db)
        return
    a = AttnArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0)
    if graph is not None:
        a.graph, a.graph_n_stride = ptr(graph), graph.shape[1] * Tk
        a.graph_q_stride = Tk if graph.shape[1] == Tq else 0
    a.key_on, a.query_on = ptr(key_on), ptr(query_on)
    a.N, a.H, a.Tq, a.Tk, a.d = N, H, Tq, Tk, d
    a.causal, a.renorm, a.engine, a.scale_d = int(causal), int(renorm), int(engine), int(scale_d)
    _check(dout, F32, "dout")
    assert dout.dim() == 2 and dout.stride(1) == 1
    a.dout, a.ld_dout = ptr(dout), dout.stride(0)
    a.dq, a.ld_dq, a.dk, a.ld_dk, a.dv, a.ld_dv = ptr(dq), dq.stride(0), ptr(dk), dk.stride(0), ptr(dv), dv.stride(0)
    for nm, t in (("dbq", dbq), ("dbk", dbk), ("dbv", dbv)):
        _check(t, F32, nm)
        assert t is None or (t.is_contiguous() and t.numel() == H * d)
    a.dbq, a.dbk, a.dbv = ptr(dbq), ptr(dbk), ptr(dbv)
    if graph is not None:
        _set_graph_bits(a, graph_bits, N, Tq, Tk)
    if engine == 0 and stats is not None and fwd_out is not None:
        # single-pass tcgen05 backward: the forward's row statistics and output replace the recomputation of max / Z / R / U
        _check(stats, F32, "stats")
        _check(fwd_out, F32, "fwd_out")
        assert stats.is_contiguous() and stats.numel() == H * N * Tq * 4
        assert fwd_out.dim() == 2 and fwd_out.stride(1) == 1 and fwd_out.shape[0] == N * Tq and fwd_out.shape[1] >= H * d
        a.stats, a.out, a.ldo = ptr(stats), ptr(fwd_out), fwd_out.stride(0)
    if engine == 1 and Tq > 1:
        scratch = torch.empty(2, H * N, Tq, Tk, device=q.device, dtype=F32)
        a.scratch = ptr(scratch)
    call("savqa_graph_attn_bwd", C.byref(a))


def answer_loss(lc: Tensor, lv: Tensor, ls: Tensor, answer: Tensor, epsilon: float, grad_scale: float, want_grads: bool):
    """main_itp_ddp_tar_super_node.py:335-345.  Returns (loss [1], (d_concat, d_vis, d_syb) | None)."""
    for nm, t in (("logits_concat", lc), ("logits_vis", lv), ("logits_syb", ls)):
        _check(t, F32, nm)
        assert t.is_contiguous() and t.shape == lc.shape
    _check(answer, torch.int64, "answer")
    B, ncls = lc.shape
    loss = torch.empty(1, device=lc.device, dtype=F32)
    grads = tuple(torch.empty_like(lc) for _ in range(3)) if want_grads else (None, None, None)
    call("savqa_answer_loss", ptr(lc), ptr(lv), ptr(ls), ptr(answer.contiguous()), B, ncls, float(epsilon), float(grad_scale), ptr(loss),
         ptr(grads[0]), ptr(grads[1]), ptr(grads[2]))
    return loss, (grads if want_grads else None)


def adam_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, lr: float, beta1: float, beta2: float, eps: float,
              step: int, dyn: Optional[Tensor] = None, param_bf16: Optional[Tensor] = None) -> None:
    """Fused Adam over flat buffers; param_bf16 (optional) receives the bf16 mirror of the updated parameters."""
    for nm, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _check(t, F32, nm)
        assert t.is_contiguous() and t.numel() == param.numel()
    _check(param_bf16, BF16, "param_bf16")
    assert param_bf16 is None or (param_bf16.is_contiguous() and param_bf16.numel() == param.numel())
    call("savqa_adam_step", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1, beta2, eps, int(step), ptr(dyn),
         ptr(param_bf16))


def adam_advance(dyn: Tensor, lr: float, beta1: float, beta2: float) -> None:
    """dyn = {lr / (1 - beta1^step), sqrt(1 - beta2^step), step} with step = dyn[2] + 1, computed ON the device."""
    _check(dyn, F32, "dyn")
    assert dyn.numel() == 3 and dyn.is_contiguous()
    call("savqa_adam_advance", ptr(dyn), float(lr), float(beta1), float(beta2))


def adam_rows(param: Tensor, grad: Optional[Tensor], exp_avg: Tensor, exp_avg_sq: Tensor, row_stamp: Tensor, idx: Optional[Tensor], lr: float,
              beta1: float, beta2: float, eps: float, step: int, dyn: Optional[Tensor] = None, apply: bool = True) -> None:
    """Row-sparse Adam with dense-Adam semantics (savqa_adam_rows): rows named in idx (each once) first replay the zero-gradient
    steps they missed, then -- apply=True -- take this step's update (their gradient rows are zeroed).  apply=False is the
    catch-up alone (through step - 1); idx=None then covers the whole table."""
    for nm, t in (("param", param), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        _check(t, F32, nm)
    q48 = grad is not None and grad.dtype == torch.int64  # the fixed-point accumulator of scatter_add_rows
    if not q48:
        _check(grad, F32, "grad")
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t is None or (t.is_contiguous() and t.shape == param.shape)
    _check(row_stamp, torch.int32, "row_stamp")
    _check(idx, torch.int64, "idx")
    assert idx is not None or not apply
    idx = idx.contiguous() if idx is not None else None
    call("savqa_adam_rows", ptr(param), ptr(grad), int(q48), ptr(exp_avg), ptr(exp_avg_sq), ptr(row_stamp), param.shape[0], param.shape[1], ptr(idx),
         idx.numel() if idx is not None else 0, lr, beta1, beta2, eps, int(step), ptr(dyn), int(bool(apply)))
