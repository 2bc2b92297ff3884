"""torch.autograd.Function wrappers that sequence the sm_100a kernels for the forward AND backward of the
reference's transformer primitives (modules.py).  Every FLOP on these paths runs in libsavqa_b200.so; torch
only allocates buffers and threads the autograd graph.

Numerics contract: residual stream, LayerNorm, softmax and all reductions in fp32; only the
MMA operands (activations / weights / probabilities) are bf16, accumulation is fp32.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .ops import BF16, F32, pad8

Tensor = torch.Tensor

#: training under CUDA-graph replay must re-stage the bf16 weights inside the graph every step
FORCE_RESTAGE = False
#: bumped by the trainer after it updates parameters with its own kernels (which do not touch torch's version counters)
WEIGHT_EPOCH = 0
#: attention forward engine: 0 = tcgen05/TMEM kernel, 1 = CUDA-core verification kernel
ATTN_ENGINE = 0
#: bound weight packs launch their weight-gradient GEMMs on a second stream (see WeightPack.weight_grad)
WGRAD_SIDE_STREAM = True
#: SMs the persistent GEMMs on the side streams may occupy (0 = all): the rest stays free for the main streams' kernels, so the
#: decoder's chain of small launches does not queue behind whole GEMMs (savqa_set_gemm_sm_limit)
SIDE_GEMM_SMS = int(os.environ.get("SAVQA_SIDE_SMS", "132"))
_WGRAD_STREAMS = {}  # compute stream -> its weight-gradient stream
_WGRAD_DIRTY = []    # weight-gradient streams with work launched since the last join


def wgrad_stream_of(cur: "torch.cuda.Stream") -> "torch.cuda.Stream":
    key = (cur.device_index, cur.cuda_stream)
    s = _WGRAD_STREAMS.get(key)
    if s is None:
        s = _WGRAD_STREAMS[key] = torch.cuda.Stream(device=cur.device)
    return s


def join_wgrad_streams() -> None:
    """Makes the current stream wait for every weight-gradient stream (before the gradients are reduced / applied)."""
    cur = torch.cuda.current_stream()
    for s in _WGRAD_DIRTY:  # only streams used since the last join (inside a graph capture: only captured ones)
        cur.wait_stream(s)
    _WGRAD_DIRTY.clear()


def tc_attention_fits(d: int, Tk: int) -> bool:
    """Whether the tcgen05 attention kernel (csrc/attn_tcgen05.cu) takes this shape: head size 64 / 128 and Q, K, V
    and the bf16 probability tile of one (sample, head) inside one CTA's 227 KB of shared memory, scores in 512 TMEM
    columns.  Other shapes run on the CUDA-core kernel (engine 1)."""
    if d == 32:
        d = 64  # zero-padded to 64-wide tiles by ops.graph_attention_fwd (score scale stays 1/sqrt(32))
    if d not in (64, 128) or Tk > 512:
        return False
    dch = d // 64
    tk16 = (Tk + 15) // 16 * 16
    box = min(tk16, 256)
    kv_rows = (tk16 + box - 1) // box * box
    qk = max(dch * 16384 + dch * kv_rows * 128, ((Tk + 63) // 64) * 16384)  # the P tile overlays Q and K
    smem = 1024 + qk + dch * kv_rows * 128 + Tk * 4 + 16 + 128 * ((Tk + 31) // 32) * 4
    return smem + 64 <= 227 * 1024


class Side:
    """bf16 copy and row mask of an activation produced by one of our kernels, handed to the next module so that
    it does not have to re-read the fp32 tensor (valid while the fp32 tensor is not modified in place)."""
    __slots__ = ("bf16", "on", "version", "data_ptr")

    def __init__(self, t: Tensor, bf16: Optional[Tensor], on: Optional[Tensor]):
        self.bf16, self.on, self.version, self.data_ptr = bf16, on, t._version, t.data_ptr()

    @staticmethod
    def of(t: Tensor) -> Optional["Side"]:
        s = getattr(t, "_savqa_side", None)
        if s is not None and s.version == t._version and s.data_ptr == t.data_ptr():
            return s
        return None


class WeightPack:
    """bf16 MMA-operand staging of one or more nn.Linear layers that read the same input, concatenated along the
    output dimension: W [sum N, pad8(K)] (forward reads it K-major, dgrad reads the same bytes MN-major), fp32 bias [sum N].

    Two modes:
      * unbound (plain autograd use, tests): re-staged by cast kernels whenever a parameter's storage or version
        counter changes; backward returns the parameter gradients to autograd;
      * bound (train.EncoderTrainer): `w` / `bias` are views of the trainer's flat bf16 mirror / flat fp32 parameter
        buffer (the optimizer kernel keeps the mirror current, so nothing is staged per step) and `gw` / `gb` are views
        of the flat gradient buffer that the backward kernels accumulate into directly (autograd gets None)."""

    def __init__(self):
        self.key = None
        self.w: Optional[Tensor] = None
        self.bias: Optional[Tensor] = None
        self.gw: Optional[Tensor] = None
        self.gb: Optional[Tensor] = None
        self.bound = False

    def bind(self, w_bf16: Tensor, bias_f32: Tensor, gw: Tensor, gb: Tensor) -> "WeightPack":
        assert w_bf16.dtype == BF16 and w_bf16.dim() == 2 and w_bf16.shape[1] % 8 == 0 and w_bf16.data_ptr() % 16 == 0
        assert gw.shape == w_bf16.shape and gb.shape == bias_f32.shape == (w_bf16.shape[0],)
        self.w, self.bias, self.gw, self.gb, self.bound = w_bf16, bias_f32, gw, gb, True
        return self

    def unbind(self) -> None:
        self.w = self.bias = self.gw = self.gb = self.key = None
        self.bound = False

    def refresh(self, weights: Sequence[Tensor], biases: Sequence[Tensor]) -> "WeightPack":
        if self.bound:
            return self
        key = (WEIGHT_EPOCH,) + tuple((w.data_ptr(), w._version, b.data_ptr(), b._version) for w, b in zip(weights, biases))
        if key == self.key and not FORCE_RESTAGE:
            return self
        K = weights[0].shape[1]
        ntot = sum(w.shape[0] for w in weights)
        dev = weights[0].device
        if self.w is None or self.w.shape != (ntot, pad8(K)) or self.w.device != dev:
            self.w = torch.empty(ntot, pad8(K), device=dev, dtype=BF16)
            self.bias = torch.empty(ntot, device=dev, dtype=F32)
        r = 0
        for w, b in zip(weights, biases):
            n = w.shape[0]
            ops.cast_bf16(w.detach(), out=self.w[r:r + n], pad_to=pad8(K))
            self.bias[r:r + n].copy_(b.detach())
            r += n
        self.key = key
        return self

    # ---- backward helpers -------------------------------------------------------------------------------
    def bias_grad_buffer(self, n: int, dev) -> Tensor:
        """fp32 [n] that the producer of dY accumulates the bias gradient into (the flat-gradient view when bound)."""
        return self.gb if self.bound else torch.zeros(n, device=dev, dtype=F32)

    def weight_grad(self, dyb: Tensor, xb: Tensor, n_out: int, k_in: int, bias_grad: Optional[Tensor] = None) -> Tensor:
        """dW[n_out, k_in] (+)= dY^T X on the tensor cores, into the flat-gradient view when bound.  `bias_grad` (fp32 [n_out],
        optional) += the column sums of dY on the same stream: a bias gradient that no epilogue upstream could produce cheaply
        (the K = 512 dgrad GEMM that writes dY is epilogue-bound: the in-epilogue column sums cost it 40 % -- profiles/r1_06).

        Bound packs: nothing downstream in the backward pass reads dW (only the optimizer does), so the GEMM goes to a second
        stream and leaves the dgrad chain -- which is a chain of launch-latency-bound M = B kernels in the decoder -- alone.
        The trainer joins the streams before the all-reduce / Adam (join_wgrad_streams)."""
        if self.bound:
            dW = self.gw if self.gw.shape[1] == k_in else self.gw[:, :k_in]
            if WGRAD_SIDE_STREAM and dyb.is_cuda:
                cur = torch.cuda.current_stream()
                side = wgrad_stream_of(cur)
                side.wait_stream(cur)
                if side not in _WGRAD_DIRTY:
                    _WGRAD_DIRTY.append(side)
                with torch.cuda.stream(side), ops.gemm_sm_limit(SIDE_GEMM_SMS):
                    ops.wgrad(dyb, xb, n_out, k_in, dW)
                    if bias_grad is not None:
                        ops.colsum_bf16(dyb[:, :n_out], bias_grad)
                # the operands are temporaries of the caller: keep their memory from being re-used before the side stream is done
                dyb.record_stream(side)
                xb.record_stream(side)
                return dW
        else:
            dW = torch.zeros(n_out, k_in, device=dyb.device, dtype=F32)
        ops.wgrad(dyb, xb, n_out, k_in, dW)
        if bias_grad is not None:
            ops.colsum_bf16(dyb[:, :n_out], bias_grad)
        return dW

    def out(self, t: Optional[Tensor]) -> Optional[Tensor]:
        """What backward hands to autograd for a parameter gradient: nothing when it already sits in the flat buffer."""
        return None if self.bound else t


def dgrad(dyb: Tensor, pack: WeightPack, M: int, k_in: int, n_out: int, **epilogue) -> None:
    """dX[M, k_in] = dY[M, n_out] W[n_out, k_in]: the weight staging is read MN-major (no transposed copy exists)."""
    ops.gemm(dyb, pack.w, M, k_in, n_out, b_mn=True, **epilogue)


def bind_linear(pack: WeightPack, lin, fv) -> None:
    """Binds the pack of one nn.Linear to the trainer's flat buffers (fv) when the layer is in them and its input width is
    a multiple of 8 (16-byte TMA row pitch in the bf16 mirror); fv None unbinds."""
    if fv is None:
        pack.unbind()
        return
    n, k = lin.weight.shape
    if k % 8 or not fv.has([lin.weight, lin.bias]):
        return
    pack.bind(fv.bf16([lin.weight]).view(n, k), fv.param([lin.bias]), fv.grad([lin.weight]).view(n, k), fv.grad([lin.bias]))


class FlatViews:
    """Index of the trainer's flat buffers: parameter -> slice.  `param` / `grad` are fp32 [n], `mirror` is the bf16 copy
    of `param` that the optimizer kernel maintains."""

    def __init__(self, param: Tensor, grad: Tensor, mirror: Tensor, offsets: dict):
        self._param, self._grad, self._mirror, self._off = param, grad, mirror, offsets  # offsets: id(p) -> (start, numel)

    def has(self, params) -> bool:
        if not all(id(p) in self._off for p in params):
            return False
        o = self._off[id(params[0])][0]
        for p in params:  # the group must be contiguous, in this order
            if self._off[id(p)][0] != o:
                return False
            o += self._off[id(p)][1]
        return True

    def _span(self, params):
        assert self.has(params), "parameters are not contiguous in the flat buffers"
        lo = self._off[id(params[0])][0]
        return lo, lo + sum(self._off[id(p)][1] for p in params)

    def param(self, params) -> Tensor:
        lo, hi = self._span(params)
        return self._param[lo:hi]

    def grad(self, params) -> Tensor:
        lo, hi = self._span(params)
        return self._grad[lo:hi]

    def bf16(self, params) -> Tensor:
        lo, hi = self._span(params)
        assert lo % 8 == 0
        return self._mirror[lo:hi]


class NormSink:
    """Gradient sinks of one layer_normalization (gamma, beta): views of the trainer's flat gradient buffer when bound."""

    def __init__(self):
        self.dgamma: Optional[Tensor] = None
        self.dbeta: Optional[Tensor] = None
        self.bound = False

    def bind(self, dgamma: Tensor, dbeta: Tensor) -> None:
        self.dgamma, self.dbeta, self.bound = dgamma, dbeta, True

    def unbind(self) -> None:
        self.dgamma = self.dbeta = None
        self.bound = False

    def buffers(self, gamma: Tensor):
        if self.bound:
            return self.dgamma, self.dbeta
        return torch.zeros_like(gamma), torch.zeros_like(gamma)

    def out(self, t: Optional[Tensor]) -> Optional[Tensor]:
        return None if self.bound else t


_NO_SINK = NormSink()


def _grouped_rows(t: Tensor):
    """A [B, R, W] slice of a larger row-major tensor (stride (S * W, W, 1), S >= R: what torch.cat's backward hands out) seen as the
    2-D matrix it lives in: (matrix [(B - 1) S + R, W], R, S) for the kernels that address rows in equally spaced groups -- or None."""
    if t.dim() != 3 or not t.is_cuda:
        return None
    B, R, W = t.shape
    sb, sr, sw = t.stride()
    if sw != 1 or sr != W or sb % W or sb < R * W or (t.data_ptr() % 16) or t.dtype not in (F32, BF16):
        return None
    S = sb // W
    return torch.as_strided(t, ((B - 1) * S + R, W), (W, 1)), R, S


def _as_bf16_rows(x: Tensor, M: int, K: int) -> Tensor:
    """[M, pad8(K)] bf16 staging of an activation (cast kernel for fp32, view for bf16)."""
    if x.dtype == BF16:
        x2 = x.reshape(M, K)
        if x2.stride(1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()
        assert K % 8 == 0, "bf16 activations must have a multiple-of-8 width"
        return x2
    return ops.cast_bf16(x.reshape(M, K).contiguous() if not x.is_contiguous() else x.reshape(M, K))


# ======================================================================================================
# nn.Linear (+ReLU) (+ broadcast row table)  -- AttModel_x3.py:42-44, 97-101; classifier heads :482-500
# ======================================================================================================
class LinearFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, rowtab, pack: WeightPack, relu: bool, out_bf16: bool, period: int):
        K = x.shape[-1]
        N = weight.shape[0]
        M = x.numel() // K
        pack.refresh([weight], [bias])
        xb = _as_bf16_rows(x, M, K)
        yb = torch.empty(M, N, device=x.device, dtype=BF16) if (out_bf16 or relu) and N % 8 == 0 else None
        y32 = None if out_bf16 else torch.empty(M, N, device=x.device, dtype=F32)
        if (out_bf16 or relu) and yb is None:
            raise ValueError("savqa_b200: bf16 / ReLU outputs need an output width that is a multiple of 8")
        rt = rowtab.detach()[:period] if rowtab is not None else None
        ops.gemm(xb, pack.w, M, N, K, bias=pack.bias, rowtab=rt, rowtab_period=period if rt is not None else 0, relu=relu,
                 out_f32=y32, out_bf16=yb)
        ctx.pack, ctx.relu, ctx.dims, ctx.period = pack, relu, (M, N, K), period
        ctx.x_dtype, ctx.x_shape = x.dtype, x.shape
        ctx.rowtab_shape = None if rowtab is None else rowtab.shape
        ctx.save_for_backward(xb, yb if relu else None)
        out = yb if out_bf16 else y32
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xb, act = ctx.saved_tensors
        pack: WeightPack = ctx.pack
        M, N, K = ctx.dims
        grp = _grouped_rows(dy) if (ctx.relu and not dy.is_contiguous()) else None
        dy2 = dy.reshape(M, N) if grp is None else None
        if dy2 is not None and not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if grp is not None:  # a slice of a concatenated gradient (the question rows of d x_in): gated where it lies, no copy first
            dyb = ops.relu_gate_bf16(grp[0], act, group_rows=grp[1], group_stride=grp[2])
        elif ctx.relu:
            dyb = ops.relu_gate_bf16(dy2, act)
        elif dy2.dtype == BF16:
            dyb = dy2
        else:
            dyb = ops.cast_bf16(dy2)
        dev = dy.device
        dW = db = dx = drowtab = None
        if ctx.needs_input_grad[1]:
            dW = pack.out(pack.weight_grad(dyb, xb, N, K))
        if ctx.needs_input_grad[2]:
            db = pack.bias_grad_buffer(N, dev)
            ops.colsum_bf16(dyb[:, :N], db)
            db = pack.out(db)
        if ctx.needs_input_grad[0]:
            if ctx.x_dtype == BF16:
                dx = torch.empty(M, K, device=dev, dtype=BF16)
                dgrad(dyb, pack, M, K, N, out_bf16=dx)
            else:
                dx = torch.empty(M, K, device=dev, dtype=F32)
                dgrad(dyb, pack, M, K, N, out_f32=dx)
            dx = dx.reshape(ctx.x_shape)
        if ctx.rowtab_shape is not None and ctx.needs_input_grad[3]:
            # d rowtab[t] = sum_b dy[b, t]; tiny fp32 reduction over the batch (positional table, AttModel_x3.py:100)
            drowtab = torch.zeros(ctx.rowtab_shape, device=dev, dtype=F32)
            drowtab[:ctx.period] = dy2.float().reshape(-1, ctx.period, N).sum(0)
            if ctx.period == ctx.rowtab_shape[0]:
                drowtab[-1] = 0  # `embedding(zeros_pad=False)` passes padding_idx=-1: the LAST row is frozen (modules.py:34-41)
        return dx, dW, db, drowtab, None, None, None, None


# ======================================================================================================
# embedding gather  -- nn.Embedding (AttModel_x3.py:96) and modules.embedding (modules.py:32-46)
# ======================================================================================================
class RowGradLog:
    """Row-sparse gradient hand-off for the 407000 x 300 word tables (train.py): instead of materialising a dense
    488 MB gradient per step, EmbeddingFn.backward records (row ids, row gradients) here and returns no table gradient;
    the trainer exchanges / scatters / applies them (savqa_scatter_add_rows + savqa_adam_rows)."""

    def __init__(self):
        self.pending = []  # list of (flat_idx int64 [n], rows fp32 [n, width], scale, skip_row)
        #: trainer hooks: before_read(ids) runs right before a gather reads the table (deferred Adam: the rows are caught up on the
        #: stream that is about to read them); on_grad() right after a gradient list was appended (the update can start while the rest
        #: of the backward pass runs)
        self.before_read = None
        self.on_grad = None

    def clear(self):
        self.pending.clear()


class EmbeddingFn(Function):
    @staticmethod
    def forward(ctx, idx, table, scale: float, skip_row: int, rowlog: Optional[RowGradLog] = None):
        flat = idx.reshape(-1)
        if rowlog is not None and rowlog.before_read is not None:
            rowlog.before_read(flat)
        out, _ = ops.gather_rows(table.detach(), flat, scale=scale, want_f32=True)
        ctx.save_for_backward(flat)
        ctx.table_shape, ctx.scale, ctx.skip_row, ctx.rowlog = table.shape, scale, skip_row, rowlog
        return out.reshape(*idx.shape, table.shape[1])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (flat,) = ctx.saved_tensors
        d2 = dy.reshape(flat.numel(), -1)
        d2 = d2 if d2.is_contiguous() else d2.contiguous()
        if ctx.rowlog is not None:
            ctx.rowlog.pending.append((flat, d2, ctx.scale, ctx.skip_row))
            if ctx.rowlog.on_grad is not None:
                ctx.rowlog.on_grad()
            return None, None, None, None, None
        dtable = torch.zeros(ctx.table_shape, device=dy.device, dtype=F32)  # dense, like the reference
        ops.scatter_add_rows(dtable, flat, d2, scale=ctx.scale, skip_row=ctx.skip_row)
        return None, dtable, None, None, None


# ======================================================================================================
# layer_normalization  -- modules.py:49-65
# ======================================================================================================
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, sink: Optional[NormSink] = None):
        xc = x if x.is_contiguous() else x.contiguous()
        y, _, yb, on = ops.layernorm_fwd(xc, None, gamma.detach(), beta.detach(), eps, save_pre=False, want_bf16=True, want_on=True)
        ctx.save_for_backward(xc, gamma)
        ctx.eps = eps
        ctx.sink = sink or _NO_SINK
        ctx.mark_non_differentiable(yb, on)
        ctx.set_materialize_grads(False)  # no zero-fill launches for the gradients of the non-differentiable outputs
        return y, yb, on

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dyb, _don):
        if dy is None:
            return None, None, None, None, None
        x, gamma = ctx.saved_tensors
        dg, db = ctx.sink.buffers(gamma)
        dx, _ = ops.layernorm_bwd(dy.contiguous(), x, gamma.detach(), ctx.eps, dg, db)
        return dx, ctx.sink.out(dg), ctx.sink.out(db), None, None


# ======================================================================================================
# attention modules  -- modules.py:119-207 (renorm 0), :210-311 (renorm 1), :314-403 (renorm 2)
# ======================================================================================================
#: set by train.EncoderTrainer when gradients are reduced across ranks: an object with `.ready(key)` that is told, from inside
#: the backward pass, that every gradient of the parameter bucket `key` has been launched (train.GradReducer)
GRAD_REDUCER = None


class BucketMarkFn(Function):
    """Identity whose backward reports that the parameters used DOWNSTREAM of this point (one encoder block, the heads ...)
    have all had their gradient kernels launched, so that their all-reduce can start while the rest of the backward runs."""

    @staticmethod
    def forward(ctx, x, key):
        ctx.key = key
        ctx.set_materialize_grads(False)
        y = x.view_as(x)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        if GRAD_REDUCER is not None:
            GRAD_REDUCER.ready(ctx.key)
        return dy, None


def bucket_mark(x: Tensor, key) -> Tensor:
    """Marks `x` (keeping its bf16 / row-mask side information) when a gradient reducer is active; otherwise returns it as is."""
    if GRAD_REDUCER is None or not (x.requires_grad and torch.is_grad_enabled()):
        return x
    side = Side.of(x)
    y = BucketMarkFn.apply(x, key)
    if side is not None:
        y._savqa_side = Side(y, side.bf16, side.on)
    return y


class MemoryHolder:
    """Side channel between the decoder's cross-attention layers and the encoder output they all read (`memory`,
    AttModel_x3.py:148-152).  Their K/V projections of `memory` -- big GEMMs that depend on nothing in the decoder -- and the
    matching dgrad / wgrad GEMMs run on a second stream next to the decoder's chain of launch-latency-bound M = B kernels;
    the six layers' contributions to d(memory) accumulate in `dmem`, which MemoryJoinFn hands to autograd in one piece."""

    def __init__(self):
        self.ready: Optional["torch.cuda.Event"] = None   # memory (and its bf16 copy) are complete on the compute stream
        self.dmem: Optional[Tensor] = None                # fp32 [N*T, C], accumulated on the side stream
        self.side: Optional["torch.cuda.Stream"] = None
        # fused form (the branch model's [2 L C, C] K/V block is bound): ONE projection GEMM for all L layers ...
        self.kv_all: Optional[Tensor] = None              # bf16 [N*T, 2 L C]: layer i reads columns [2 C i, 2 C (i + 1))
        self.kv_done: Optional["torch.cuda.Event"] = None   # the whole block is projected
        self.kv_done0: Optional["torch.cuda.Event"] = None  # layer 0's slice is (it is projected first, by a launch of its own: the
        #                                                     decoder's first cross-attention then waits ~20 us instead of ~100)
        self.dmem_rest: Optional[Tensor] = None             # d(memory) of layers 1..L-1, launched under the decoder's last layer (backward)
        self.rest_done: Optional["torch.cuda.Event"] = None
        self.pack_all: Optional[WeightPack] = None
        self.mem_bf16: Optional[Tensor] = None
        self.dkv_all: Optional[Tensor] = None             # ... and one dgrad (K = 2 L C) + one wgrad once every layer has written its slice


class MemoryJoinFn(Function):
    """Identity on `memory` whose backward returns the d(memory) that the cross-attention layers accumulated on the side
    stream (they return no gradient for it themselves)."""

    @staticmethod
    def forward(ctx, x, holder: MemoryHolder):
        ctx.holder = holder
        ctx.set_materialize_grads(False)
        return x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        h: MemoryHolder = ctx.holder
        g = dy
        if h.dkv_all is not None:
            # every cross-attention layer has written its slice of dkv_all: d(memory) = dKV_all W_all in one GEMM (K = 2 L C) on
            # this stream (the encoder's backward is next), the weight gradient of the whole block on the side stream
            pk, dkv = h.pack_all, h.dkv_all
            Mk, n = dkv.shape
            C = pk.w.shape[1]
            d = torch.empty(Mk, C, device=dkv.device, dtype=F32)
            res = None if g is None else g.reshape(Mk, C)
            if h.dmem_rest is not None:
                # layers 1..L-1 were contracted on the side stream while the decoder's backward was still in its last layer
                # (DecoderFn.backward): only layer 0's K = 2C slice is left on the critical path
                torch.cuda.current_stream().wait_event(h.rest_done)
                if res is not None:
                    h.dmem_rest.add_(res)
                ops.gemm(dkv[:, :2 * C], pk.w[:2 * C], Mk, C, 2 * C, b_mn=True, res=h.dmem_rest, out_f32=d)
                h.dmem_rest = None
            else:
                ops.gemm(dkv, pk.w, Mk, C, n, b_mn=True, res=res, out_f32=d)
            pk.weight_grad(dkv, h.mem_bf16, n, C)
            g = d.reshape(ctx_shape(h))
            h.dkv_all = h.kv_all = None
        if GRAD_REDUCER is not None and getattr(h, "bucket", None) is not None:
            GRAD_REDUCER.ready(h.bucket)  # the decoder's backward is complete
        if h.dmem is not None:
            torch.cuda.current_stream().wait_stream(h.side)
            d = h.dmem.reshape(ctx_shape(h))
            g = d if g is None else g + d
            h.dmem = None
        return g, None


def ctx_shape(h: MemoryHolder):
    return h.shape


def _same(a: Tensor, b: Tensor) -> bool:
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a._version == b._version)


class GraphAttentionFn(Function):
    """y = LN( merge_heads( W' V ) + queries ),  W from softmax(QK^T/sqrt d) re-weighted by the graph (see ops)."""

    @staticmethod
    def forward(ctx, queries, keys, values, graph, Wq, bq, Wk, bk, Wv, bv, gamma, beta, q_bf16, q_on, k_bf16, k_on, cfg):
        H, causal, renorm, want_att, packs, eps = cfg["heads"], cfg["causal"], cfg["renorm"], cfg["return_att"], cfg["packs"], cfg["eps"]
        N, Tq, C = queries.shape
        Tk = keys.shape[1]
        d = C // H
        Mq, Mk = N * Tq, N * Tk
        same_qk = _same(queries, keys)
        same_kv = _same(keys, values)
        xq = queries if queries.is_contiguous() else queries.contiguous()

        # ---- padding masks from the RAW inputs + bf16 staging (modules.py:257, 289) ----
        if q_bf16 is None or q_on is None:
            q_on, q_bf16 = ops.row_nonzero(xq.reshape(Mq, C))
        q_bf16 = q_bf16.reshape(Mq, C)
        if same_qk:
            k_on, k_bf16 = q_on, q_bf16
        elif k_bf16 is None or k_on is None:
            k_on, k_bf16 = ops.row_nonzero(keys.reshape(Mk, C) if keys.is_contiguous() else keys.contiguous().reshape(Mk, C))
        k_bf16 = k_bf16.reshape(Mk, C)
        if same_kv:
            v_bf16 = k_bf16
        else:
            v_bf16 = ops.cast_bf16(values.reshape(Mk, C) if values.is_contiguous() else values.contiguous().reshape(Mk, C))

        # ---- projections: Linear + ReLU, fused along N when the inputs coincide (modules.py:241-243) ----
        dev = queries.device
        holder = None
        fused_kv = False
        if same_qk and same_kv:
            pk = packs["qkv"].refresh([Wq, Wk, Wv], [bq, bk, bv])
            qkv = torch.empty(Mq, 3 * C, device=dev, dtype=BF16)
            ops.gemm(q_bf16, pk.w, Mq, 3 * C, C, bias=pk.bias, relu=True, out_bf16=qkv)
            q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
            mode = 0
        else:
            pq = packs["q"].refresh([Wq], [bq])
            q = torch.empty(Mq, C, device=dev, dtype=BF16)
            ops.gemm(q_bf16, pq.w, Mq, C, C, bias=pq.bias, relu=True, out_bf16=q)
            if same_kv:
                pkv = packs["kv"].refresh([Wk, Wv], [bk, bv])
                kv = torch.empty(Mk, 2 * C, device=dev, dtype=BF16)
                holder = cfg.get("kv_holder")
                if holder is not None and not (pkv.bound and WGRAD_SIDE_STREAM and holder.ready is not None):
                    holder = None
                if holder is not None and holder.kv_all is not None:
                    # projected for all layers at once when the encoder finished (AttModel_x3._Branch._join_memory)
                    C2 = 2 * C
                    i0 = cfg.get("kv_index", 0) * C2
                    torch.cuda.current_stream().wait_event(holder.kv_done)
                    kv = holder.kv_all[:, i0:i0 + C2]
                    fused_kv = True
                elif holder is not None:
                    # depends only on `memory`: off the decoder's kernel chain, onto the side stream
                    cur = torch.cuda.current_stream()
                    holder.side.wait_event(holder.ready)
                    with torch.cuda.stream(holder.side), ops.gemm_sm_limit(SIDE_GEMM_SMS):
                        ops.gemm(k_bf16, pkv.w, Mk, 2 * C, C, bias=pkv.bias, relu=True, out_bf16=kv)
                    kv.record_stream(holder.side)
                    cur.wait_stream(holder.side)
                else:
                    ops.gemm(k_bf16, pkv.w, Mk, 2 * C, C, bias=pkv.bias, relu=True, out_bf16=kv)
                k, v = kv[:, :C], kv[:, C:]
                mode = 1
            else:
                pk_, pv_ = packs["k"].refresh([Wk], [bk]), packs["v"].refresh([Wv], [bv])
                k = torch.empty(Mk, C, device=dev, dtype=BF16)
                v = torch.empty(Mk, C, device=dev, dtype=BF16)
                ops.gemm(k_bf16, pk_.w, Mk, C, C, bias=pk_.bias, relu=True, out_bf16=k)
                ops.gemm(v_bf16, pv_.w, Mk, C, C, bias=pv_.bias, relu=True, out_bf16=v)
                mode = 2

        # ---- attention core ----
        g = gbits = None
        if graph is not None and renorm != 0:
            gbits = ops.graph_bits_of(graph)  # 0/1 graphs built by ops.build_masks carry their bit-packed form
            g = graph if graph.dtype == F32 else graph.float()
            g = g if g.is_contiguous() else g.contiguous()
            if g is not graph:
                gbits = None
        engine = ATTN_ENGINE if (tc_attention_fits(d, Tk) and Tq > 1) else 1  # Tq == 1: one-warp row kernel
        # training on the tensor-core engine: keep the softmax row statistics (and the core's output) for a one-pass backward
        stats = None
        if engine == 0 and any(ctx.needs_input_grad[:3]) and ops.tc_attention_bwd_fits(d, Tq, Tk):
            stats = torch.empty(H * N * Tq, 4, device=dev, dtype=F32)
        o, att = ops.graph_attention_fwd(q, k, v, g, k_on, q_on, N, H, Tq, Tk, d, causal, renorm if g is not None else 0, want_att, engine,
                                         graph_bits=gbits, stats=stats)

        # ---- residual (RAW queries) + LayerNorm (modules.py:304-307) ----
        y, pre, yb, y_on = ops.layernorm_fwd(o.reshape(N, Tq, C), xq, gamma.detach(), beta.detach(), eps, save_pre=True, want_bf16=True,
                                             want_on=True)
        ctx.cfg, ctx.mode, ctx.dims = cfg, mode, (N, Tq, Tk, C, H, d)
        ctx.renorm_eff = renorm if g is not None else 0
        ctx.gbits = gbits
        ctx.kv_holder = holder if mode == 1 else None
        ctx.kv_fused = fused_kv
        ctx.save_for_backward(q_bf16, k_bf16, v_bf16, q, k, v, g, q_on, k_on, pre, gamma, stats, o if stats is not None else None)
        outs = (y, yb, y_on) + ((att,) if want_att else ())
        ctx.mark_non_differentiable(yb, y_on, *((att,) if want_att else ()))
        ctx.set_materialize_grads(False)  # autograd would otherwise zero-fill a gradient for yb / y_on / att (two launches per module)
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, *_unused):
        if dy is None:
            return (None,) * 17
        q_bf16, k_bf16, v_bf16, q, k, v, g, q_on, k_on, pre, gamma, stats, fwd_o = ctx.saved_tensors
        cfg, mode = ctx.cfg, ctx.mode
        N, Tq, Tk, C, H, d = ctx.dims
        Mq, Mk = N * Tq, N * Tk
        packs = cfg["packs"]
        sink: NormSink = cfg.get("norm_sink") or _NO_SINK
        dev = dy.device
        need = ctx.needs_input_grad
        dgamma, dbeta = sink.buffers(gamma)
        dpre, _ = ops.layernorm_bwd(dy.contiguous(), pre, gamma.detach(), cfg["eps"], dgamma, dbeta)
        dpre2 = dpre.reshape(Mq, C)

        # attention core backward -> ReLU-gated dQ, dK, dV in the layout of the fused projection outputs, and the
        # projections' bias gradients (column sums of the gated gradients) from the same kernel
        if mode == 0:
            pqkv = packs["qkv"]
            dqkv = torch.empty(Mq, 3 * C, device=dev, dtype=BF16)
            dq, dk, dv = dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:]
            db = pqkv.bias_grad_buffer(3 * C, dev)
            dbq, dbk, dbv = db[:C], db[C:2 * C], db[2 * C:]
        elif mode == 1:
            pq, pkv = packs["q"], packs["kv"]
            dq = torch.empty(Mq, C, device=dev, dtype=BF16)
            if ctx.kv_fused:  # this layer's slice of the gradient of the fused K/V projection (MemoryJoinFn.backward consumes it)
                hold = ctx.kv_holder
                if hold.dkv_all is None:
                    hold.dkv_all = torch.empty(Mk, hold.pack_all.w.shape[0], device=dev, dtype=BF16)
                i0 = cfg.get("kv_index", 0) * 2 * C
                dkv = hold.dkv_all[:, i0:i0 + 2 * C]
            else:
                dkv = torch.empty(Mk, 2 * C, device=dev, dtype=BF16)
            dk, dv = dkv[:, :C], dkv[:, C:]
            dbq = pq.bias_grad_buffer(C, dev)
            dbkv = pkv.bias_grad_buffer(2 * C, dev)
            dbk, dbv = dbkv[:C], dbkv[C:]
        else:
            pq, pk_, pv_ = packs["q"], packs["k"], packs["v"]
            dq = torch.empty(Mq, C, device=dev, dtype=BF16)
            dk = torch.empty(Mk, C, device=dev, dtype=BF16)
            dv = torch.empty(Mk, C, device=dev, dtype=BF16)
            dbq, dbk, dbv = pq.bias_grad_buffer(C, dev), pk_.bias_grad_buffer(C, dev), pv_.bias_grad_buffer(C, dev)
        # encoder self-attention (fused QKV, Tq > 1): the bias gradient is one column-sum launch over dqkv next to the weight
        # gradient on the side stream -- inside the attention kernel the 6 warp-transpose reductions per thread were ~17 % of its
        # instructions (profiles/r1_06_attn_bwd_lines.txt); the one-query decoder kernels keep theirs
        bias_outside = mode == 0 and Tq > 1
        ops.graph_attention_bwd(q, k, v, g, k_on, q_on, N, H, Tq, Tk, d, cfg["causal"], ctx.renorm_eff, dpre2, dq, dk, dv,
                                dbq=None if bias_outside else dbq, dbk=None if bias_outside else dbk, dbv=None if bias_outside else dbv,
                                graph_bits=ctx.gbits, stats=stats, fwd_out=fwd_o)

        dxq = dxk = dxv = None
        if mode == 0:
            dW = pqkv.weight_grad(dqkv, q_bf16, 3 * C, C, bias_grad=db if bias_outside else None)
            dWq, dWk, dWv = dW[:C], dW[C:2 * C], dW[2 * C:]
            if need[0] or need[1] or need[2]:
                dxq = torch.empty(Mq, C, device=dev, dtype=F32)
                dgrad(dqkv, pqkv, Mq, C, 3 * C, res=dpre2, out_f32=dxq)  # + residual branch
                dxq = dxq.reshape(N, Tq, C)
            po = pqkv
        else:
            dWq = pq.weight_grad(dq, q_bf16, C, C)
            if need[0]:
                dxq = torch.empty(Mq, C, device=dev, dtype=F32)
                dgrad(dq, pq, Mq, C, C, res=dpre2, out_f32=dxq)
                dxq = dxq.reshape(N, Tq, C)
            if mode == 1 and ctx.kv_fused:
                dWk = dWv = None  # weight gradient and d(memory) of all layers together: MemoryJoinFn.backward
            elif mode == 1:
                dW = pkv.weight_grad(dkv, k_bf16, 2 * C, C)
                dWk, dWv = dW[:C], dW[C:]
                holder = ctx.kv_holder
                if holder is not None and (need[1] or need[2]):
                    # d(memory) += dKV W_kv on the side stream (the six layers serialise there); MemoryJoinFn returns the sum
                    cur = torch.cuda.current_stream()
                    first = holder.dmem is None
                    if first:
                        holder.dmem = torch.empty(Mk, C, device=dev, dtype=F32)
                        holder.dmem.record_stream(holder.side)
                    holder.side.wait_stream(cur)
                    if holder.side not in _WGRAD_DIRTY:
                        _WGRAD_DIRTY.append(holder.side)
                    with torch.cuda.stream(holder.side), ops.gemm_sm_limit(SIDE_GEMM_SMS):
                        ops.gemm(dkv, pkv.w, Mk, C, 2 * C, b_mn=True, out_f32=holder.dmem, accumulate=0 if first else 2)
                    dkv.record_stream(holder.side)
                elif need[1] or need[2]:
                    dxk = torch.empty(Mk, C, device=dev, dtype=F32)
                    dgrad(dkv, pkv, Mk, C, 2 * C, out_f32=dxk)
                    dxk = dxk.reshape(N, Tk, C)
            else:
                dWk = pk_.weight_grad(dk, k_bf16, C, C)
                dWv = pv_.weight_grad(dv, v_bf16, C, C)
                if need[1]:
                    dxk = torch.empty(Mk, C, device=dev, dtype=F32)
                    dgrad(dk, pk_, Mk, C, C, out_f32=dxk)
                    dxk = dxk.reshape(N, Tk, C)
                if need[2]:
                    dxv = torch.empty(Mk, C, device=dev, dtype=F32)
                    dgrad(dv, pv_, Mk, C, C, out_f32=dxv)
                    dxv = dxv.reshape(N, Tk, C)
            po = pq
        o = po.out  # all packs of one module are bound together
        # when queries/keys/values are one tensor autograd sums the three slots: hand the total to the first
        return (dxq, dxk, dxv, None, o(dWq), o(dbq), o(dWk), o(dbk), o(dWv), o(dbv), sink.out(dgamma), sink.out(dbeta),
                None, None, None, None, None)


class TokenSelfAttentionFn(Function):
    """multihead_attention (modules.py:119-207) on ONE token per sample attending to itself -- the decoder's self-attention
    (AttModel_x3.py:148): the softmax over a single key is 1 whatever Q and K are (a masked key gives the uniform row, also 1), so
        y = LN( relu(x Wv^T + bv) * query_mask + x )
    and W_q, b_q, W_k, b_k receive exactly zero gradient (as in the reference, where d softmax = p (1 - p) = 0).  One N = C GEMM
    instead of the N = 3C projection + the attention core, forward and backward: the decoder is a chain of launch-latency-bound
    kernels, every link counts."""

    @staticmethod
    def forward(ctx, x, Wq, bq, Wk, bk, Wv, bv, gamma, beta, x_bf16, x_on, cfg):
        packs, eps = cfg["packs"], cfg["eps"]
        N, T, C = x.shape
        assert T == 1
        xc = x if x.is_contiguous() else x.contiguous()
        if x_bf16 is None or x_on is None:
            x_on, x_bf16 = ops.row_nonzero(xc.reshape(N, C))
        x_bf16 = x_bf16.reshape(N, C)
        pv = packs["v"].refresh([Wv], [bv])
        vb = torch.empty(N, C, device=x.device, dtype=BF16)
        ops.gemm(x_bf16, pv.w, N, C, C, bias=pv.bias, relu=True, out_bf16=vb)
        o = torch.mul(vb, x_on.reshape(N, 1))  # fp32: the bf16 V row times the query mask, what W' V gives for a single key
        y, pre, yb, y_on = ops.layernorm_fwd(o.reshape(N, 1, C), xc, gamma.detach(), beta.detach(), eps, save_pre=True, want_bf16=True,
                                             want_on=True)
        ctx.cfg, ctx.dims = cfg, (N, C)
        ctx.save_for_backward(x_bf16, vb, x_on, pre, gamma, Wq, bq, Wk, bk)
        ctx.mark_non_differentiable(yb, y_on)
        ctx.set_materialize_grads(False)
        return y, yb, y_on

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, *_unused):
        if dy is None:
            return (None,) * 12
        x_bf16, vb, x_on, pre, gamma, Wq, bq, Wk, bk = ctx.saved_tensors
        cfg = ctx.cfg
        N, C = ctx.dims
        pv = cfg["packs"]["v"]
        sink: NormSink = cfg.get("norm_sink") or _NO_SINK
        dgamma, dbeta = sink.buffers(gamma)
        dpre, _ = ops.layernorm_bwd(dy.contiguous(), pre, gamma.detach(), cfg["eps"], dgamma, dbeta)
        dpre2 = dpre.reshape(N, C)
        dvb = ops.relu_gate_bf16(dpre2 * x_on.reshape(N, 1), vb)
        dbv = pv.bias_grad_buffer(C, dy.device)
        dWv = pv.weight_grad(dvb, x_bf16, C, C, bias_grad=dbv)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(N, C, device=dy.device, dtype=F32)
            dgrad(dvb, pv, N, C, C, res=dpre2, out_f32=dx)  # + residual branch
            dx = dx.reshape(N, 1, C)
        o = pv.out
        # unbound modules hand autograd explicit zeros for the Q / K projections, like the reference's autograd does (the trainer
        # discovers the parameters of its flat buffers by their gradients); bound ones leave the zero-filled flat gradient alone
        z = (lambda t: None) if pv.bound else torch.zeros_like
        return (dx, z(Wq), z(bq), z(Wk), z(bk), o(dWv), o(dbv), sink.out(dgamma), sink.out(dbeta), None, None, None)


# ======================================================================================================
# the decoder chain  -- AttModel_x3.py:141-154
# ======================================================================================================
#: the decoder's L x (one-token self-attention, one-query cross-attention, feedforward) chain on fused cluster GEMM + LayerNorm
#: launches (savqa_gemm_rowln); SAVQA_FUSED_DECODER=0 keeps the per-module chain (A/B comparison, fallback for other shapes)
FUSED_DECODER = os.environ.get("SAVQA_FUSED_DECODER", "1") != "0"


class DecoderFn(Function):
    """The whole decoder of one branch model as ONE autograd node.  Every layer is a chain of M = B-row kernels whose length is
    the sum of their latencies; here each Linear -> (mask) -> residual -> LayerNorm group is one cluster launch whose CTAs exchange
    the row statistics through distributed shared memory (csrc/rowln_tcgen05.cu), and in the backward pass every dgrad GEMM
    carries the LayerNorm backward (and the ReLU gate of the layer in front of it) in its epilogue:

        forward, per layer:   [Wv + mask + res + LN] -> [Wq] -> one-query attention -> res + LN -> [W1] -> [W2 + res + LN]     (6 launches, was 9)
        backward, per layer:  [W2^T, gate] -> [W1^T + res + LN'] -> attention' -> [Wq^T + res + LN' + gate] -> [Wv^T + res + LN'] (6, was 12)

    Weight gradients go to the side stream as before (WeightPack.weight_grad).  Training needs the trainer's bound packs (gradients
    are accumulated straight into the flat buffers) and the fused K/V block of the encoder output (MemoryHolder); inference takes
    any packs."""

    @staticmethod
    def forward(ctx, x0, memory, dec_mask, cfg):
        layers, H = cfg["layers"], cfg["heads"]
        B, C = x0.shape[0], x0.shape[-1]
        N, T = memory.shape[0], memory.shape[1]
        Mk, d = N * T, C // H
        dev = x0.device
        x = x0.reshape(B, C)
        x = x if x.is_contiguous() else x.contiguous()
        x_on, xb = ops.row_nonzero(x)
        ms = Side.of(memory)
        if ms is not None and ms.bf16 is not None and ms.on is not None:
            mem_b, mem_on = ms.bf16.reshape(Mk, C), ms.on
        else:
            mem_on, mem_b = ops.row_nonzero(memory.reshape(Mk, C) if memory.is_contiguous() else memory.contiguous().reshape(Mk, C))
        holder = cfg.get("kv_holder")
        fused_kv = holder is not None and holder.kv_all is not None
        g = dec_mask if dec_mask.dtype == F32 else dec_mask.float()
        g = g if g.is_contiguous() else g.contiguous()
        f32 = lambda *shape: torch.empty(*shape, device=dev, dtype=F32)    # noqa: E731
        b16 = lambda *shape: torch.empty(*shape, device=dev, dtype=BF16)   # noqa: E731
        saved = []
        for i, (sa, ca, ff) in enumerate(layers):
            n1, n2, n3 = sa.normalization, ca.normalization, ff.normalization
            pv = sa._packs["v"].refresh([sa.V_proj[0].weight], [sa.V_proj[0].bias])
            pq = ca._packs["q"].refresh([ca.Q_proj[0].weight], [ca.Q_proj[0].bias])
            p1 = ff._packs["w1"].refresh([ff.conv1[0].weight], [ff.conv1[0].bias])
            p2 = ff._packs["w2"].refresh([ff.conv2.weight], [ff.conv2.bias])
            Hd = p1.w.shape[0]
            # self-attention over one token: y1 = LN(relu(x Wv^T + bv) * query_mask + x)      (modules.py:119-207 with a single key)
            vb, pre1, y1, y1b, on1, st1 = b16(B, C), f32(B, C), f32(B, C), b16(B, C), f32(B), f32(B, 2)
            ops.gemm_rowln(xb, pv.w, B, C, C, 1, bias=pv.bias, relu=True, rowscale=x_on, res=x, gamma=n1.gamma.detach(), beta=n1.beta.detach(),
                           eps=n1.epsilon, act_bf16=vb, pre=pre1, y=y1, y_bf16=y1b, on=on1, stats=st1)
            # cross-attention: one query per sample over the encoder output                    (modules.py:236-311, Tq = 1)
            q = b16(B, C)
            ops.gemm_rowln(y1b, pq.w, B, C, C, 0, bias=pq.bias, relu=True, y_bf16=q)
            if fused_kv:
                if x0.is_cuda and i <= 1:  # layer 0's K/V slice is projected first; the other layers' in a second launch
                    ev = holder.kv_done0 if (i == 0 and holder.kv_done0 is not None) else holder.kv_done
                    torch.cuda.current_stream().wait_event(ev)
                kv = holder.kv_all[:, 2 * C * i:2 * C * (i + 1)]
            else:
                pkv = ca._packs["kv"].refresh([ca.K_proj[0].weight, ca.V_proj[0].weight], [ca.K_proj[0].bias, ca.V_proj[0].bias])
                kv = b16(Mk, 2 * C)
                ops.gemm(mem_b, pkv.w, Mk, 2 * C, C, bias=pkv.bias, relu=True, out_bf16=kv)
            k, v = kv[:, :C], kv[:, C:]
            o, _ = ops.graph_attention_fwd(q, k, v, g, mem_on, on1, N, H, 1, T, d, False, 1, False, 1)
            st2 = f32(B, 2)
            y2, pre2, y2b, on2 = ops.layernorm_fwd(o, y1, n2.gamma.detach(), n2.beta.detach(), n2.epsilon, save_pre=True, want_bf16=True,
                                                   want_on=True, stats=st2)
            # feedforward: y3 = LN(relu(y2 W1^T + b1) W2^T + b2 + y2)                           (modules.py:432-447)
            h = b16(B, Hd)
            ops.gemm_rowln(y2b, p1.w, B, Hd, C, 0, bias=p1.bias, relu=True, y_bf16=h)
            pre3, y3, y3b, on3, st3 = f32(B, C), f32(B, C), b16(B, C), f32(B), f32(B, 2)
            ops.gemm_rowln(h, p2.w, B, C, Hd, 1, bias=p2.bias, res=y2, gamma=n3.gamma.detach(), beta=n3.beta.detach(), eps=n3.epsilon,
                           pre=pre3, y=y3, y_bf16=y3b, on=on3, stats=st3)
            saved.append(dict(xb=xb, x_on=x_on, vb=vb, pre1=pre1, st1=st1, y1b=y1b, on1=on1, q=q, k=k, v=v, pre2=pre2, st2=st2, y2b=y2b,
                              h=h, pre3=pre3, st3=st3))
            x, xb, x_on = y3, y3b, on3
        ctx.cfg, ctx.saved_layers, ctx.dims = cfg, saved, (B, C, N, T, H, d)
        ctx.g, ctx.mem_on, ctx.fused_kv = g, mem_on, fused_kv
        ctx.mark_non_differentiable(xb, x_on)
        ctx.set_materialize_grads(False)
        return x.reshape(B, 1, C), xb.reshape(B, 1, C), x_on

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, *_unused):
        if dy is None:
            return None, None, None, None
        cfg, saved = ctx.cfg, ctx.saved_layers
        layers = cfg["layers"]
        B, C, N, T, H, d = ctx.dims
        Mk = N * T
        dev = dy.device
        holder = cfg.get("kv_holder")
        assert ctx.fused_kv and holder is not None, "savqa_b200: the fused decoder's backward needs the trainer's bound K/V block"
        f32 = lambda *shape: torch.empty(*shape, device=dev, dtype=F32)    # noqa: E731
        b16 = lambda *shape: torch.empty(*shape, device=dev, dtype=BF16)   # noqa: E731
        L = len(layers)
        # LayerNorm backward of the top layer's feedforward: its input gradient comes from the heads, not from one of our GEMMs
        ff = layers[L - 1][2]
        n3, p2 = ff.normalization, ff._packs["w2"]
        dg, db = n3._sink.buffers(n3.gamma)
        dy2 = dy.reshape(B, C)
        dz, dzb = ops.layernorm_bwd(dy2 if dy2.is_contiguous() else dy2.contiguous(), saved[L - 1]["pre3"], n3.gamma.detach(), n3.epsilon, dg, db,
                                    want_bf16=True, dxsum=p2.bias_grad_buffer(C, dev))
        dx0 = None
        for i in range(L - 1, -1, -1):
            s = saved[i]
            sa, ca, ff = layers[i]
            n1, n2 = sa.normalization, ca.normalization
            pv, pq, pkv, p1, p2 = sa._packs["v"], ca._packs["q"], ca._packs["kv"], ff._packs["w1"], ff._packs["w2"]
            Hd = p1.w.shape[0]
            # feedforward: dh = (dz W2) * [h > 0];  d y2 = dh W1 + dz, and straight on through the cross-attention's LayerNorm
            p2.weight_grad(dzb, s["h"], C, Hd)
            dh = b16(B, Hd)
            ops.gemm_rowln(dzb, p2.w, B, Hd, C, 0, b_mn=True, gate=s["h"], y_bf16=dh)
            p1.weight_grad(dh, s["y2b"], Hd, C, bias_grad=p1.bias_grad_buffer(Hd, dev))
            dg2, db2 = n2._sink.buffers(n2.gamma)
            dpre2 = f32(B, C)
            ops.gemm_rowln(dh, p1.w, B, C, Hd, 2, b_mn=True, res=dz, pre=s["pre2"], stats=s["st2"], gamma=n2.gamma.detach(), eps=n2.epsilon,
                           y=dpre2, dgamma=dg2, dbeta=db2)
            # cross-attention core (one query): ReLU-gated dq, this layer's slice of the fused K/V gradient, the projections' bias gradients
            if holder.dkv_all is None:
                holder.dkv_all = b16(Mk, holder.pack_all.w.shape[0])
            dkv = holder.dkv_all[:, 2 * C * i:2 * C * (i + 1)]
            dq = b16(B, C)
            dbkv = pkv.bias_grad_buffer(2 * C, dev)
            ops.graph_attention_bwd(s["q"], s["k"], s["v"], ctx.g, ctx.mem_on, s["on1"], N, H, 1, T, d, False, 1, dpre2, dq, dkv[:, :C], dkv[:, C:],
                                    dbq=pq.bias_grad_buffer(C, dev), dbk=dbkv[:C], dbv=dbkv[C:])
            pq.weight_grad(dq, s["y1b"], C, C)
            if i == 1 and L > 1 and dev.type == "cuda" and WGRAD_SIDE_STREAM:
                # every layer but 0 has written its slice of the fused K/V gradient: contract them now (K = 2 (L - 1) C) on the side
                # stream, under the rest of the chain; MemoryJoinFn.backward adds layer 0's slice
                cur = torch.cuda.current_stream()
                side = wgrad_stream_of(cur)
                side.wait_stream(cur)
                if side not in _WGRAD_DIRTY:
                    _WGRAD_DIRTY.append(side)
                pall, dall = holder.pack_all, holder.dkv_all
                with torch.cuda.stream(side), ops.gemm_sm_limit(SIDE_GEMM_SMS):
                    holder.dmem_rest = torch.empty(Mk, C, device=dev, dtype=F32)
                    ops.gemm(dall[:, 2 * C:], pall.w[2 * C:], Mk, C, dall.shape[1] - 2 * C, b_mn=True, out_f32=holder.dmem_rest)
                    holder.rest_done = torch.cuda.Event()
                    holder.rest_done.record(side)
                holder.dmem_rest.record_stream(cur)
            # d y1 = dq Wq + dpre2 -> the self-attention's LayerNorm backward -> dvb = (dpre1 * query_mask) * [vb > 0]
            dg1, db1 = n1._sink.buffers(n1.gamma)
            dpre1, dvb = f32(B, C), b16(B, C)
            ops.gemm_rowln(dq, pq.w, B, C, C, 2, b_mn=True, res=dpre2, pre=s["pre1"], stats=s["st1"], gamma=n1.gamma.detach(), eps=n1.epsilon,
                           y=dpre1, dxg_bf16=dvb, gate=s["vb"], rowscale=s["x_on"], dgamma=dg1, dbeta=db1)
            pv.weight_grad(dvb, s["xb"], C, C, bias_grad=pv.bias_grad_buffer(C, dev))
            # d x = dvb Wv + dpre1: the gradient of the layer below's output -> its feedforward LayerNorm backward in the same launch
            if i > 0:
                ffp = layers[i - 1][2]
                n3p, p2p = ffp.normalization, ffp._packs["w2"]
                dg3, db3 = n3p._sink.buffers(n3p.gamma)
                dz, dzb = f32(B, C), b16(B, C)
                ops.gemm_rowln(dvb, pv.w, B, C, C, 2, b_mn=True, res=dpre1, pre=saved[i - 1]["pre3"], stats=saved[i - 1]["st3"],
                               gamma=n3p.gamma.detach(), eps=n3p.epsilon, y=dz, y_bf16=dzb, dgamma=dg3, dbeta=db3,
                               dxsum=p2p.bias_grad_buffer(C, dev))
            else:
                dx0 = f32(B, C)
                ops.gemm_rowln(dvb, pv.w, B, C, C, 0, b_mn=True, res=dpre1, y=dx0)
        ctx.saved_layers = None
        return dx0.reshape(B, 1, C), None, None, None


# ======================================================================================================
# feedforward  -- modules.py:405-447
# ======================================================================================================
class FeedForwardFn(Function):
    """y = LN( relu(x W1^T + b1) W2^T + b2 + x )"""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, gamma, beta, x_bf16, cfg):
        packs, eps = cfg["packs"], cfg["eps"]
        C = x.shape[-1]
        Hd = W1.shape[0]
        M = x.numel() // C
        xc = x if x.is_contiguous() else x.contiguous()
        xb = x_bf16.reshape(M, C) if x_bf16 is not None else ops.cast_bf16(xc.reshape(M, C))
        p1 = packs["w1"].refresh([W1], [b1])
        p2 = packs["w2"].refresh([W2], [b2])
        dev = x.device
        h = torch.empty(M, Hd, device=dev, dtype=BF16)
        ops.gemm(xb, p1.w, M, Hd, C, bias=p1.bias, relu=True, out_bf16=h)
        z = torch.empty(M, C, device=dev, dtype=F32)
        ops.gemm(h, p2.w, M, C, Hd, bias=p2.bias, res=xc.reshape(M, C), out_f32=z)
        y, _, yb, y_on = ops.layernorm_fwd(z.reshape(x.shape), None, gamma.detach(), beta.detach(), eps, save_pre=False, want_bf16=True,
                                           want_on=True)
        ctx.cfg, ctx.dims = cfg, (M, C, Hd)
        ctx.save_for_backward(xb, h, z, gamma)
        ctx.mark_non_differentiable(yb, y_on)
        ctx.set_materialize_grads(False)
        return y, yb, y_on

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, *_unused):
        if dy is None:
            return (None,) * 9
        xb, h, z, gamma = ctx.saved_tensors
        cfg = ctx.cfg
        M, C, Hd = ctx.dims
        p1, p2 = cfg["packs"]["w1"], cfg["packs"]["w2"]
        sink: NormSink = cfg.get("norm_sink") or _NO_SINK
        dev = dy.device
        dgamma, dbeta = sink.buffers(gamma)
        # conv2: z = h W2^T + b2 + x.  db2 = column sums of dz come out of the LayerNorm backward kernel
        db2 = p2.bias_grad_buffer(C, dev)
        dz, dzb = ops.layernorm_bwd(dy.contiguous().reshape(M, C), z, gamma.detach(), cfg["eps"], dgamma, dbeta, want_bf16=True, dxsum=db2)
        dW2 = p2.weight_grad(dzb, h, C, Hd)
        # conv1: h = relu(x W1^T + b1).  ReLU backward fused in the dgrad epilogue; db1 = column sums of dh next to the weight
        # gradient (off the dgrad chain)
        db1 = p1.bias_grad_buffer(Hd, dev)
        dh = torch.empty(M, Hd, device=dev, dtype=BF16)
        dgrad(dzb, p2, M, Hd, C, gate=h, out_bf16=dh)
        dW1 = p1.weight_grad(dh, xb, Hd, C, bias_grad=db1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, C, device=dev, dtype=F32)
            dgrad(dh, p1, M, C, Hd, res=dz, out_f32=dx)  # + residual branch
            dx = dx.reshape(dy.shape)
        return dx, p1.out(dW1), p1.out(db1), p2.out(dW2), p2.out(db2), sink.out(dgamma), sink.out(dbeta), None, None


# ======================================================================================================
# MIL_NCE  -- AttModel_x3.py:285-443 (only_obj=True)
# ======================================================================================================
class MilNceFn(Function):
    """(new_macro_ipt [B,M,2048] bf16, mil_nce_obj) = MIL_NCE(vis_fea, word ids ...): one gather for the three id lists, the four
    small Linear+ReLU layers on the tensor-core GEMM, the per-object score / logsumexp / softmax-refine / scatter stage in one
    kernel (savqa_mil_nce_fwd), forward and backward.  marco_mlp and the macro-node word rows receive no gradient (`.detach()`
    at AttModel_x3.py:354), exactly like the reference."""

    @staticmethod
    def forward(ctx, vis_fea, table, Wm, bm, Ws, bs, Wv, bv, Wi, bi, macro_ipt, loc, pos_ids, neg_ids, mask, packs, rowlog):
        B, V = vis_fea.shape[0], vis_fea.shape[1]
        M, topN, h, E, Fd = macro_ipt.shape[1], pos_ids.shape[2], Ws.shape[0], table.shape[1], vis_fea.shape[-1]
        if h % 8 or topN > 8:
            raise ValueError("savqa_b200: MIL_NCE needs hidden_size_mil % 8 == 0 and topN <= 8")
        dev = vis_fea.device
        vis_b = _as_bf16_rows(vis_fea, B * V, Fd)
        ids_pn = torch.cat([pos_ids.reshape(-1), neg_ids.reshape(-1)])
        ids_all = torch.cat([macro_ipt.reshape(-1), ids_pn])
        if rowlog is not None and rowlog.before_read is not None:
            rowlog.before_read(ids_all)
        _, x = ops.gather_rows(table.detach(), ids_all, want_f32=False, want_bf16=True)
        xm, xpn = x[:B * M], x[B * M:]
        pm, ps = packs["marco"].refresh([Wm], [bm]), packs["syb"].refresh([Ws], [bs])
        pv, pi = packs["vis"].refresh([Wv], [bv]), packs["ipt"].refresh([Wi], [bi])
        n = B * V * topN
        nodes = torch.empty(B * M, h, device=dev, dtype=BF16)
        ops.gemm(xm, pm.w, B * M, h, E, bias=pm.bias, relu=True, out_bf16=nodes)            # :352 (detached at :354)
        pn_h = torch.empty(2 * n, h, device=dev, dtype=BF16)
        ops.gemm(xpn, ps.w, 2 * n, h, E, bias=ps.bias, relu=True, out_bf16=pn_h)            # :356-359
        vis_h = torch.empty(B * V, h, device=dev, dtype=BF16)
        ops.gemm(vis_b, pv.w, B * V, h, Fd, bias=pv.bias, relu=True, out_bf16=vis_h)        # :361
        mask_i = mask.reshape(-1).to(torch.int32).contiguous()
        loc_l = loc.reshape(-1).to(torch.int64).contiguous()
        raw, obj = ops.mil_nce_fwd(pn_h, vis_h, mask_i, loc_l, nodes, B, V, M, topN, h)     # :365-379
        out = torch.empty(B * M, Wi.shape[0], device=dev, dtype=BF16)
        ops.gemm(nodes, pi.w, B * M, Wi.shape[0], h, bias=pi.bias, relu=True, out_bf16=out)  # :441
        ctx.packs, ctx.rowlog, ctx.dims, ctx.table_shape = packs, rowlog, (B, V, M, topN, h, E, Fd, Wi.shape[0]), table.shape
        ctx.save_for_backward(xpn, vis_b, pn_h, vis_h, nodes, out, raw, mask_i, loc_l, ids_pn)
        ctx.set_materialize_grads(False)
        return out.view(B, M, Wi.shape[0]), obj.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out, d_obj):
        xpn, vis_b, pn_h, vis_h, nodes, out, raw, mask_i, loc_l, ids_pn = ctx.saved_tensors
        B, V, M, topN, h, E, Fd, Fo = ctx.dims
        packs = ctx.packs
        ps, pv, pi = packs["syb"], packs["vis"], packs["ipt"]
        dev = out.device
        if d_out is None and d_obj is None:
            return (None,) * 17
        d_nodes = dWi = dbi = None
        if d_out is not None:
            # d_out usually is the node-row part of d x_in [B, T, 2048] (the symbolic branch concatenates the question rows behind
            # the nodes, AttModel_x3.py:218-219): gate it where it lies instead of copying it out first
            grp = _grouped_rows(d_out)
            if grp is not None:
                dpre = ops.relu_gate_bf16(grp[0], out, group_rows=grp[1], group_stride=grp[2])
            else:
                dpre = ops.relu_gate_bf16(d_out.reshape(B * M, Fo).contiguous(), out)
            dbi = pi.bias_grad_buffer(Fo, dev)
            dWi = pi.weight_grad(dpre, nodes, Fo, h, bias_grad=dbi)
            d_nodes = torch.empty(B * M, h, device=dev, dtype=F32)
            dgrad(dpre, pi, B * M, h, Fo, out_f32=d_nodes)
        dobj = d_obj.reshape(1).float().contiguous() if d_obj is not None else None
        d_pn, d_vis = ops.mil_nce_bwd(pn_h, vis_h, mask_i, loc_l, raw, d_nodes, dobj, B, V, M, topN, h)
        dbs, dbv = ps.bias_grad_buffer(h, dev), pv.bias_grad_buffer(h, dev)
        dWs = ps.weight_grad(d_pn, xpn, h, E, bias_grad=dbs)
        dWv = pv.weight_grad(d_vis, vis_b, h, Fd, bias_grad=dbv)
        dtable = None
        if ctx.needs_input_grad[1]:
            dx = torch.empty(2 * B * V * topN, E, device=dev, dtype=F32)
            dgrad(d_pn, ps, 2 * B * V * topN, E, h, out_f32=dx)
            if ctx.rowlog is not None:
                ctx.rowlog.pending.append((ids_pn, dx, 1.0, -1))
                if ctx.rowlog.on_grad is not None:
                    ctx.rowlog.on_grad()
            else:
                dtable = torch.zeros(ctx.table_shape, device=dev, dtype=F32)  # dense, like the reference
                ops.scatter_add_rows(dtable, ids_pn, dx)
        return (None, dtable, None, None, ps.out(dWs), ps.out(dbs), pv.out(dWv), pv.out(dbv), pi.out(dWi) if dWi is not None else None,
                pi.out(dbi) if dbi is not None else None, None, None, None, None, None, None, None)


# ======================================================================================================
# three-head label-smoothed loss  -- main_itp_ddp_tar_super_node.py:335-345
# ======================================================================================================
class AnswerLossFn(Function):
    @staticmethod
    def forward(ctx, logits_concat, logits_vis, logits_syb, answer, epsilon: float):
        lc, lv, ls = (t.contiguous() for t in (logits_concat, logits_vis, logits_syb))
        loss, grads = ops.answer_loss(lc, lv, ls, answer, epsilon, 1.0, want_grads=True)
        ctx.save_for_backward(*grads)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        dc, dv, ds = ctx.saved_tensors
        return dc * dloss, dv * dloss, ds * dloss, None, None
