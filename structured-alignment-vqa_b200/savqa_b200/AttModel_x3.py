"""Drop-in for the reference's `models/AttModel_x3.py` (`--model_v 3`): same classes, constructor arguments,
forward signatures, output shapes and state_dict keys; the graph-guided encoder / decoder path runs on the
sm_100a kernels (modules.py of this package).

    AttModel_vis_grid  <- AttModel_x3.py:20-156      AttModel_syb <- :158-282      AttModel <- :471-542
    MIL_NCE            <- :285-443  (gather + small GEMMs + one fused score kernel, forward and backward;
                                     only_obj=True only)
    CompactBilinearPooling <- :444-469 (parameter holder only: torch.rfft no longer exists, `mcb=True` raises)

Deliberate differences: no hard-coded `.cuda()` (inputs decide the device, which must be a B200); the per-sample
Python loop with a host sync per sample that builds the masks (:110-116) is one kernel launch.
"""
from __future__ import annotations

import math

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from . import ops
from .functional import WeightPack
from .modules import *  # noqa: F401,F403  (the reference does `from modules import *`)
from .modules import embedding, feedforward, label_smoothing, multihead_attention, new_multihead_attention

PAD = 400000
UNK = 400001
END = 400003
INVALID = 400003
VIS_PAD = -1
LOC_PAD = -1
VOCAB_ROWS = 407000  # hard-coded in the reference (AttModel_x3.py:36, 168, 293)
_SIDE_STREAMS = {}   # device -> second stream of AttModel.encoder_step
_HEAD_STREAMS = {}   # device -> the two extra streams of AttModel.answer_logits
_HEADS_PARALLEL = os.environ.get("SAVQA_HEADS_PARALLEL", "1") != "0"
#: SMs the persistent GEMMs of the (visual, symbolic) branch stream may take: a static split in proportion to the branches' work (T = 56 vs
#: 128 tokens per sample) lets one branch's GEMM compute while the other's pays its fixed launch / pipeline-fill / tail cost, instead of the
#: two taking turns on all 148 SMs (7.70 -> 7.41 ms/step measured; "0,0" = no split)
_BRANCH_SMS = tuple(int(x) for x in os.environ.get("SAVQA_BRANCH_SMS", "56,92").split(","))
if not any(_BRANCH_SMS):
    _BRANCH_SMS = None


class _CastBf16(torch.autograd.Function):
    """fp32 -> bf16 staging of the region / symbolic node features (the A operand of syb_mlp2)."""

    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.contiguous().reshape(-1, x.shape[-1])).reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        return dy.float()


def _bucket_blocks() -> int:
    from . import train
    return train.BUCKET_BLOCKS


def _word_table(glove) -> nn.Embedding:
    table = torch.empty(VOCAB_ROWS, 300)
    nn.init.xavier_normal_(table)
    n = glove.vectors.shape[0]
    table[:n, :] = glove.vectors
    return nn.Embedding.from_pretrained(table, freeze=False)


def _linear(x, lin: nn.Linear, pack: WeightPack, relu=False, out_bf16=False, rowtab=None, period=0):
    return Fn.LinearFn.apply(x, lin.weight, lin.bias, rowtab, pack, relu, out_bf16, period)


class _Branch(nn.Module):
    """Shared forward of the two branch models (they differ in the top-left block of `graph` and in table sizes)."""

    def _cross_layers(self):
        return [getattr(self, 'dec_vanilla_attention_%d' % i) for i in range(self.num_blocks)]

    def _savqa_groups(self):
        """[Wk_0; Wv_0; ...; Wk_{L-1}; Wv_{L-1}] of the decoder's cross-attention layers as ONE [2 L C, C] block (and their biases
        as one [2 L C] vector): every layer projects the same encoder output (AttModel_x3.py:148-152), so the L K/V projections
        are one GEMM with N = 2 L C forward, one dgrad with K = 2 L C and one wgrad backward (functional.MemoryHolder)."""
        ws, bs = [], []
        for m in self._cross_layers():
            ws += [m.K_proj[0].weight, m.V_proj[0].weight]
            bs += [m.K_proj[0].bias, m.V_proj[0].bias]
        return [ws, bs] if ws else []

    def _savqa_bind(self, fv):
        Fn.bind_linear(self._pk["mlp"], self.syb_mlp[0], fv)    # K = 300: stays on the per-step staging path
        Fn.bind_linear(self._pk["mlp2"], self.syb_mlp2, fv)
        pk = self._pk.setdefault("kv_all", WeightPack())
        if fv is None:
            pk.unbind()
            return
        g = self._savqa_groups()
        C = self.hidden_size
        if g and C % 8 == 0 and fv.has(g[0]) and fv.has(g[1]):
            n = 2 * self.num_blocks * C
            pk.bind(fv.bf16(g[0]).view(n, C), fv.param(g[1]), fv.grad(g[0]).view(n, C), fv.grad(g[1]))

    def _masks_aside(self, first_mask, q_mask, q_graph, first_graph, dec_on, compact=None):
        """Mask construction (AttModel_x3.py:103-122 / :229-247) depends on nothing the input MLPs produce: on a GPU it runs on a
        helper stream next to them; the caller joins with the returned event before the first attention.  `compact` =
        (first_len, q_len, q_graph_bits, first_graph_bits, V, Q): the loader's compact hand-off (collate.py) instead of dense planes."""
        def build():
            if compact is not None:
                return ops.build_masks_compact(*compact, dec_on)
            return ops.build_masks(first_mask, q_mask, q_graph, first_graph, dec_on)  # (+ the bit-packed forms of the two graphs)
        probe = compact[0] if compact is not None else first_mask
        if not probe.is_cuda:
            return build(), None
        cur = torch.cuda.current_stream()
        aside = Fn.wgrad_stream_of(cur)
        aside.wait_stream(cur)
        with torch.cuda.stream(aside):
            masks = build()
            done = torch.cuda.Event()
            done.record(aside)
        for m in masks:
            m.record_stream(cur)
            bits = ops.graph_bits_of(m)
            if bits is not None:
                bits.record_stream(cur)
        return masks, done

    def _input_stage(self, first_ipt, q_ids, pos_table, pos_dropout_p):
        B = first_ipt.shape[0]
        q = Fn.EmbeddingFn.apply(q_ids, self.syb_emb.weight, 1.0, -1, getattr(self.syb_emb, "_savqa_rowlog", None))  # :96 / :216
        qh = _linear(q, self.syb_mlp[0], self._pk["mlp"], relu=True, out_bf16=True)        # :97  [B,Q,2048] bf16
        fb = first_ipt if first_ipt.dtype == torch.bfloat16 else _CastBf16.apply(first_ipt)
        x_in = torch.cat([fb, qh], dim=1)                                                  # :98  [B,T,2048] bf16
        T = x_in.shape[1]
        if T > pos_table.shape[0]:
            raise IndexError("index out of range in self")  # positional table too short, as F.embedding would report
        if self.training and pos_dropout_p > 0:
            x = _linear(x_in, self.syb_mlp2, self._pk["mlp2"])
            pos = pos_table[:T]
            if T == pos_table.shape[0]:  # padding_idx=-1: the table's last row is frozen (modules.py:34-41)
                pos = torch.cat([pos[:-1], pos[-1:].detach()], 0)
            # Sequential(embedding, Dropout) drops the gathered [B,T,C] tensor: an independent mask per sample (:100-101, vis only)
            x = x + F.dropout(pos.unsqueeze(0).expand(B, T, -1), pos_dropout_p, True)
        else:
            x = _linear(x_in, self.syb_mlp2, self._pk["mlp2"], rowtab=pos_table, period=T)  # :99-101 fused
        return self.enc_dropout(x)                                                         # :102

    def _encode_decode(self, x, graph_diag, graph, dec_mask):
        B, C = x.shape[0], x.shape[-1]
        for i in range(self.num_blocks):
            # blocks 0,1: graph_diag; the reference aliases graph_cross and graph, so 2.. all see `graph` (:118-139)
            g = graph_diag if i < 2 else graph
            x = Fn.bucket_mark(x, (id(self), "enc", i // _bucket_blocks()))  # multi-GPU: block i's gradients are final once backward passes here
            x = getattr(self, 'enc_self_attention_%d' % i)(x, x, x, g)
            x = getattr(self, 'enc_feed_forward_%d' % i)(x)
        memory = self._join_memory(x)
        # decoder input: class token 2 through dec_emb (scaled by sqrt C) + position 0 (:141-147); same for every sample
        dec = self.dec_emb.lookup_table[2] * (C ** 0.5) + self.dec_positional_encoding.lookup_table[0]
        dec = dec.reshape(1, 1, C).expand(B, 1, C).contiguous()
        dec = self.dec_dropout(dec)
        if self._fused_decoder_ok(dec, memory):
            layers = [(getattr(self, 'dec_self_attention_%d' % i), getattr(self, 'dec_vanilla_attention_%d' % i),
                       getattr(self, 'dec_feed_forward_%d' % i)) for i in range(self.num_blocks)]
            y, yb, on = Fn.DecoderFn.apply(dec, memory, dec_mask, dict(layers=layers, heads=self.num_heads,
                                                                       kv_holder=getattr(memory, "_savqa_kv_holder", None)))
            y._savqa_side = Fn.Side(y, yb, on)
            return y
        for i in range(self.num_blocks):
            dec = getattr(self, 'dec_self_attention_%d' % i)(dec, dec, dec)
            dec = getattr(self, 'dec_vanilla_attention_%d' % i)(dec, memory, memory, dec_mask)
            dec = getattr(self, 'dec_feed_forward_%d' % i)(dec)
        return dec

    def _fused_decoder_ok(self, dec, memory) -> bool:
        """Whether the decoder runs as functional.DecoderFn (fused cluster GEMM + LayerNorm launches): a shape the row-wise epilogue
        takes, no active dropout, and -- when gradients are needed -- the trainer's bound packs with the fused K/V block."""
        C = self.hidden_size
        if not (Fn.FUSED_DECODER and self.num_blocks > 0 and dec.is_cuda == memory.is_cuda and ops.rowln_fits(dec.shape[0], C, C)
                and C % self.num_heads == 0 and (4 * C) % 64 == 0):
            return False
        if not dec.is_cuda and not getattr(ops.gemm_rowln, "__module__", "").endswith("fake_ops"):
            return False
        # (dec_dropout acts on the start token BEFORE the decoder stack, AttModel_x3.py:147: it does not stand in the way of the fusion)
        ff0 = self.dec_feed_forward_0
        if ff0.num_units[0] % 64 or ff0.num_units[1] != C:
            return False
        if torch.is_grad_enabled() and (dec.requires_grad or memory.requires_grad):
            holder = getattr(memory, "_savqa_kv_holder", None)
            bound = all(getattr(self, 'dec_self_attention_%d' % i)._packs["v"].bound and getattr(self, 'dec_vanilla_attention_%d' % i)._packs["q"].bound
                        and getattr(self, 'dec_vanilla_attention_%d' % i)._packs["kv"].bound
                        and getattr(self, 'dec_feed_forward_%d' % i)._packs["w1"].bound and getattr(self, 'dec_feed_forward_%d' % i)._packs["w2"].bound
                        and getattr(self, 'dec_self_attention_%d' % i).normalization._sink.bound
                        for i in range(self.num_blocks))
            return bool(bound and holder is not None and holder.kv_all is not None)
        return True

    def _join_memory(self, x):
        """Training on a B200 with bound weight packs: routes the cross-attention K/V projections of the encoder output (and
        their backward) to a second stream, see functional.MemoryHolder.  Anything else: the plain tensor."""
        pk = self.dec_vanilla_attention_0._packs["kv"] if self.num_blocks > 0 else None
        if pk is None or not (pk.bound and Fn.WGRAD_SIDE_STREAM and x.is_cuda and x.requires_grad and torch.is_grad_enabled()):
            return x
        side_info = Fn.Side.of(x)
        h = Fn.MemoryHolder()
        cur = torch.cuda.current_stream()
        h.side = Fn.wgrad_stream_of(cur)
        h.shape = x.shape
        h.bucket = (id(self), "dec")
        memory = Fn.MemoryJoinFn.apply(x, h)
        if side_info is not None:
            memory._savqa_side = Fn.Side(memory, side_info.bf16, side_info.on)
        h.ready = torch.cuda.Event()
        h.ready.record(cur)
        memory._savqa_kv_holder = h
        pall = self._pk.get("kv_all")
        if pall is not None and pall.bound and side_info is not None and side_info.bf16 is not None:
            # all L cross-attention layers' K/V projections of the encoder output in one GEMM, on the side stream
            C = x.shape[-1]
            Mk, n = x.numel() // C, pall.w.shape[0]
            mem_bf16 = side_info.bf16.reshape(Mk, C)
            kv_all = torch.empty(Mk, n, device=x.device, dtype=torch.bfloat16)
            h.side.wait_event(h.ready)
            with torch.cuda.stream(h.side), Fn.ops.gemm_sm_limit(Fn.SIDE_GEMM_SMS):
                if self.num_blocks > 1 and Fn.FUSED_DECODER:
                    # layer 0's slice first, in a launch of its own: the decoder's first cross-attention waits for 1/L of the work
                    Fn.ops.gemm(mem_bf16, pall.w[:2 * C], Mk, 2 * C, C, bias=pall.bias[:2 * C], relu=True, out_bf16=kv_all[:, :2 * C])
                    h.kv_done0 = torch.cuda.Event()
                    h.kv_done0.record(h.side)
                    Fn.ops.gemm(mem_bf16, pall.w[2 * C:], Mk, n - 2 * C, C, bias=pall.bias[2 * C:], relu=True, out_bf16=kv_all[:, 2 * C:])
                else:
                    Fn.ops.gemm(mem_bf16, pall.w, Mk, n, C, bias=pall.bias, relu=True, out_bf16=kv_all)
            kv_all.record_stream(h.side)
            mem_bf16.record_stream(h.side)
            h.kv_done = torch.cuda.Event()
            h.kv_done.record(h.side)
            h.kv_all, h.pack_all, h.mem_bf16 = kv_all, pall, mem_bf16
        return memory

    def _add_blocks(self, which, hidden_size):
        for i in range(self.num_blocks):
            if which == "enc":
                setattr(self, 'enc_self_attention_%d' % i, new_multihead_attention(num_units=hidden_size, num_heads=self.num_heads,
                                                                                    dropout_rate=0, causality=False))
                setattr(self, 'enc_feed_forward_%d' % i, feedforward(hidden_size, [4 * hidden_size, hidden_size]))
            else:
                setattr(self, 'dec_self_attention_%d' % i, multihead_attention(num_units=hidden_size, num_heads=self.num_heads,
                                                                                dropout_rate=0, causality=True))
                setattr(self, 'dec_vanilla_attention_%d' % i, new_multihead_attention(num_units=hidden_size, num_heads=self.num_heads,
                                                                                       dropout_rate=0, causality=False))
                m = getattr(self, 'dec_vanilla_attention_%d' % i)
                m._kv_external, m._kv_index = True, i
                setattr(self, 'dec_feed_forward_%d' % i, feedforward(hidden_size, [4 * hidden_size, hidden_size]))


class AttModel_vis_grid(_Branch):
    def __init__(self, glove, hidden_size, maxlen, maxlen_q, num_blocks, num_heads, dropout_rate, maxlen_v, num_classes):
        super().__init__()
        self.hidden_size = hidden_size
        self.num_blocks = num_blocks
        self.num_heads = num_heads
        self.maxlen_q = maxlen_q
        self.maxlen = maxlen
        self.dropout_rate = dropout_rate
        self.enc_dropout = nn.Dropout(dropout_rate)
        self.dec_dropout = nn.Dropout(dropout_rate)
        self.syb_emb = _word_table(glove)
        self.syb_mlp = nn.Sequential(nn.Linear(300, 2048), nn.ReLU(inplace=True))
        self.syb_mlp2 = nn.Linear(2048, hidden_size)
        # allocated but never used by the reference's forward; kept so that checkpoints load with strict=True
        self.v_mlp = nn.Sequential(nn.Linear(2048, hidden_size), nn.ReLU(inplace=True), nn.Linear(hidden_size, hidden_size))
        self.v_positional_encoding = nn.Sequential(embedding(maxlen_v, hidden_size, zeros_pad=False, scale=False), nn.Dropout(dropout_rate))
        self.input_proj = nn.Linear(2048, hidden_size)
        self._add_blocks("enc", hidden_size)
        self.q_mlp = nn.Sequential(nn.Linear(300, hidden_size), nn.ReLU(inplace=True), nn.Linear(hidden_size, hidden_size))
        self.q_positional_encoding = nn.Sequential(embedding(maxlen_q, hidden_size, zeros_pad=False, scale=False), nn.Dropout(dropout_rate))
        self.syb_positional_encoding = nn.Sequential(embedding(maxlen, hidden_size, zeros_pad=False, scale=False), nn.Dropout(dropout_rate))
        self.dec_emb = embedding(num_classes, hidden_size, scale=True)
        self.dec_positional_encoding = embedding(maxlen, hidden_size, zeros_pad=False, scale=False)
        self._add_blocks("dec", hidden_size)
        self._pk = {"mlp": WeightPack(), "mlp2": WeightPack()}

    def forward(self, vis_fea, vis_mask, q_fea, q_graph, q_mask, decMask):
        if vis_fea.dim() == 4:  # bs x gridx x gridy x fea_size
            vis_fea = vis_fea.reshape(-1, vis_fea.size(1) * vis_fea.size(2), vis_fea.size(3))
        (graph_diag, graph, dec_mask), done = self._masks_aside(vis_mask, q_mask, q_graph, None, decMask != False)  # noqa: E712
        x = self._input_stage(vis_fea, q_fea, self.syb_positional_encoding[0].lookup_table, self.dropout_rate)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return self._encode_decode(x, graph_diag, graph, dec_mask)

    def forward_compact(self, vis_fea, vis_len, q_fea, q_graph_bits, q_len, decMask):
        """forward() on the loader's compact hand-off (collate.py): `vis_len` / `q_len` int32 [B] for the prefix-block masks, the
        question graph bit-packed.  Same result as forward() on the dense planes, bit for bit."""
        if vis_fea.dim() == 4:
            vis_fea = vis_fea.reshape(-1, vis_fea.size(1) * vis_fea.size(2), vis_fea.size(3))
        compact = (vis_len, q_len, q_graph_bits, None, vis_fea.shape[1], q_fea.shape[1])
        (graph_diag, graph, dec_mask), done = self._masks_aside(None, None, None, None, decMask != False, compact)  # noqa: E712
        x = self._input_stage(vis_fea, q_fea, self.syb_positional_encoding[0].lookup_table, self.dropout_rate)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return self._encode_decode(x, graph_diag, graph, dec_mask)


class AttModel_syb(_Branch):
    def __init__(self, glove, hidden_size, maxlen, maxlen_q, num_blocks, num_heads, dropout_rate, num_classes):
        super().__init__()
        self.num_blocks = num_blocks
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.dropout_rate = dropout_rate
        self.maxlen = maxlen
        self.maxlen_q = maxlen_q
        self.syb_emb = _word_table(glove)
        self.syb_mlp = nn.Sequential(nn.Linear(300, 2048), nn.ReLU(inplace=True))
        self.syb_mlp2 = nn.Linear(2048, hidden_size)
        self.enc_dropout = nn.Dropout(dropout_rate)
        self.syb_positional_encoding = embedding(maxlen + maxlen_q, hidden_size, zeros_pad=False, scale=False)
        self.q_mlp = nn.Sequential(nn.Linear(300, hidden_size), nn.Linear(hidden_size, hidden_size))
        self.q_positional_encoding = nn.Sequential(embedding(maxlen_q, hidden_size, zeros_pad=False, scale=False), nn.Dropout(dropout_rate))
        self.dec_emb = embedding(num_classes, hidden_size, scale=True)
        self.dec_positional_encoding = embedding(maxlen + maxlen_q, hidden_size, zeros_pad=False, scale=False)
        self.dec_dropout = nn.Dropout(dropout_rate)
        self._add_blocks("dec", hidden_size)
        self._add_blocks("enc", hidden_size)
        self._pk = {"mlp": WeightPack(), "mlp2": WeightPack()}

    def forward(self, syb_ipt, syb_mask, syb_graph, q_fea, q_graph, q_mask, decMask):
        (graph_diag, graph, dec_mask), done = self._masks_aside(syb_mask, q_mask, q_graph, syb_graph, decMask != False)  # noqa: E712
        x = self._input_stage(syb_ipt, q_fea, self.syb_positional_encoding.lookup_table, 0.0)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return self._encode_decode(x, graph_diag, graph, dec_mask)

    def forward_compact(self, syb_ipt, macro_len, macro_graph_bits, q_fea, q_graph_bits, q_len, decMask):
        """forward() on the loader's compact hand-off (collate.py); same result as forward() on the dense planes, bit for bit."""
        compact = (macro_len, q_len, q_graph_bits, macro_graph_bits, syb_ipt.shape[1], q_fea.shape[1])
        (graph_diag, graph, dec_mask), done = self._masks_aside(None, None, None, None, decMask != False, compact)  # noqa: E712
        x = self._input_stage(syb_ipt, q_fea, self.syb_positional_encoding.lookup_table, 0.0)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return self._encode_decode(x, graph_diag, graph, dec_mask)


class MIL_NCE(nn.Module):
    """Object-word alignment head (AttModel_x3.py:285-443) on the sm_100a kernels (functional.MilNceFn): produces `new_macro_ipt`
    [B,M,2048] for the symbolic branch and the object MIL-NCE term.  Only the production `only_obj=True` configuration is
    implemented (the relation branch holds a Python loop over relations, :421-436, and is outside the scope)."""

    def __init__(self, glove, hidden_size, dropout_rate, num_relations, only_obj):
        super().__init__()
        self.only_obj = only_obj
        self.dropout_rate = dropout_rate
        self.num_relations = num_relations
        self.hidden_size = hidden_size
        self.R = nn.Parameter(torch.empty(num_relations, hidden_size, hidden_size))
        nn.init.xavier_normal_(self.R)
        self.syb_emb = _word_table(glove)
        self.marco_mlp = nn.Sequential(nn.Linear(300, hidden_size), nn.ReLU(inplace=True))
        self.syb_mlp = nn.Sequential(nn.Linear(300, hidden_size), nn.ReLU(inplace=True))
        self.vis_mlp = nn.Sequential(nn.Linear(2048, hidden_size), nn.ReLU(inplace=True))
        self.rel_mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(inplace=True), nn.Linear(hidden_size, 1))
        self.softmax = nn.Softmax(dim=2)
        self.softmax_bilinear = nn.Softmax(dim=0)
        self.bilinear = nn.Bilinear(hidden_size, hidden_size, num_relations, bias=False)
        self.ipt_mlp = nn.Sequential(nn.Linear(hidden_size, 2048), nn.ReLU(inplace=True))
        self._pk = {k: WeightPack() for k in ("marco", "syb", "vis", "ipt")}
        self._bf16_out = False  # AttModel.forward keeps new_macro_ipt in bf16 (the symbolic branch's first GEMM casts it anyway)

    def _savqa_bind(self, fv):
        Fn.bind_linear(self._pk["vis"], self.vis_mlp[0], fv)   # K = 2048
        Fn.bind_linear(self._pk["ipt"], self.ipt_mlp[0], fv)   # K = hidden_size_mil; the K = 300 layers stay on the staging path

    def forward(self, vis_fea, macro_ipt, macro_obj_loc, micro_positive_obj, micro_negative_obj, micro_obj_mask,
                micro_positive_rel, micro_negative_rel, micro_positive_rel_loc, micro_negative_rel_loc):
        if not self.only_obj:
            raise NotImplementedError("savqa_b200: the relation branch of MIL_NCE (only_obj=False) is outside the scope of this build")
        out, obj = Fn.MilNceFn.apply(vis_fea, self.syb_emb.weight, self.marco_mlp[0].weight, self.marco_mlp[0].bias,
                                     self.syb_mlp[0].weight, self.syb_mlp[0].bias, self.vis_mlp[0].weight, self.vis_mlp[0].bias,
                                     self.ipt_mlp[0].weight, self.ipt_mlp[0].bias, macro_ipt, macro_obj_loc, micro_positive_obj,
                                     micro_negative_obj, micro_obj_mask, self._pk, getattr(self.syb_emb, "_savqa_rowlog", None))
        return (out if self._bf16_out else out.float()), obj, 0


class CompactBilinearPooling(nn.Module):
    """Parameter holder for state_dict compatibility (AttModel_x3.py:444-469); the reference's forward needs the
    removed torch.rfft and is off by default (`--mcb` False)."""

    def __init__(self, input_dims, output_dim):
        super().__init__()
        self.output_dim = output_dim

        def sketch():
            h = torch.randint(output_dim, size=(input_dims,))
            s = (2 * torch.randint(2, size=(input_dims,)) - 1).float()
            m = torch.zeros(input_dims, output_dim)
            m[torch.arange(input_dims), h] = s
            return nn.Parameter(m, requires_grad=False)

        self.sketch1 = sketch()
        self.sketch2 = sketch()

    def forward(self, x1, x2):
        raise NotImplementedError("compact bilinear pooling relies on torch.rfft, removed from PyTorch; run with mcb=False")


def _head(cin, hidden, ncls, p):
    return nn.Sequential(nn.Linear(cin, hidden), nn.ReLU(), nn.Dropout(p, inplace=True), nn.Linear(hidden, ncls))


class AttModel(nn.Module):
    def __init__(self, glove, hidden_size, hidden_size_mil, num_classes, maxlen_q, maxlen, maxlen_v, num_blocks, num_heads,
                 dropout_rate, dropout_rate_mcb, num_relations, only_obj):
        super().__init__()
        self.only_obj = only_obj
        self.att_vis_grid = AttModel_vis_grid(glove, hidden_size, maxlen, maxlen_q, num_blocks, num_heads, dropout_rate, maxlen_v, num_classes)
        self.att_syb = AttModel_syb(glove, hidden_size, maxlen, maxlen_q, num_blocks, num_heads, dropout_rate, num_classes)
        self.MIL_NCE = MIL_NCE(glove, hidden_size_mil, dropout_rate, num_relations, self.only_obj)
        self.MIL_NCE._bf16_out = True
        self.cls = _head(hidden_size * 2, hidden_size, num_classes, dropout_rate)
        self.cls_vis = _head(hidden_size, hidden_size, num_classes, dropout_rate)
        self.cls_syb = _head(hidden_size, hidden_size, num_classes, dropout_rate)
        self.mcb_out = 16000
        self.mcb = CompactBilinearPooling(hidden_size, self.mcb_out)
        self.mcb_dropout = nn.Dropout(dropout_rate_mcb)
        self.cls_mcb = _head(self.mcb_out, hidden_size, num_classes, dropout_rate)
        self.label_smoothing = label_smoothing()
        self._pk = {k: (WeightPack(), WeightPack()) for k in ("cls", "cls_vis", "cls_syb")}
        self.concurrent_branches = True  # encoder_step runs att_vis_grid and att_syb on two streams

    def _savqa_bind(self, fv):
        for name, (p0, p3) in self._pk.items():
            head = getattr(self, name)
            Fn.bind_linear(p0, head[0], fv)
            Fn.bind_linear(p3, head[3], fv)

    def _classify(self, name, fea):
        head = getattr(self, name)
        p0, p3 = self._pk[name]
        h = _linear(fea, head[0], p0, relu=True, out_bf16=True)
        if self.training and head[2].p > 0:
            h = F.dropout(h, head[2].p, True)
        return _linear(h, head[3], p3)

    def answer_logits(self, fea_vis_grid, fea_syb):
        """Classifier heads of AttModel_x3.py:531-541 on the two [B,1,C] decoder outputs."""
        if not (fea_vis_grid.is_cuda and _HEADS_PARALLEL):
            logits_vis = self._classify("cls_vis", fea_vis_grid).squeeze(1)
            logits_syb = self._classify("cls_syb", fea_syb).squeeze(1)
            fea = torch.cat((fea_syb.squeeze(1), fea_vis_grid.squeeze(1)), 1)
            logits_concat = self._classify("cls", fea)
            return logits_concat, logits_vis, logits_syb
        # The three heads are independent chains of two M = B GEMMs (the second one 1845 classes wide): on a GPU they run side by
        # side on three streams, forward and -- through autograd -- backward, instead of back to back while the SMs idle.
        main = torch.cuda.current_stream()
        hs = _HEAD_STREAMS.get(fea_vis_grid.device)
        if hs is None:
            hs = _HEAD_STREAMS[fea_vis_grid.device] = (torch.cuda.Stream(device=fea_vis_grid.device), torch.cuda.Stream(device=fea_vis_grid.device))
        for st, fea_ in zip(hs, (fea_vis_grid, fea_syb)):
            st.wait_stream(main)
            fea_.record_stream(st)
        with torch.cuda.stream(hs[0]):
            logits_vis = self._classify("cls_vis", fea_vis_grid).squeeze(1)
        with torch.cuda.stream(hs[1]):
            logits_syb = self._classify("cls_syb", fea_syb).squeeze(1)
        fea = torch.cat((fea_syb.squeeze(1), fea_vis_grid.squeeze(1)), 1)
        logits_concat = self._classify("cls", fea)
        for st, lg in zip(hs, (logits_vis, logits_syb)):
            main.wait_stream(st)
            lg.record_stream(main)
        return logits_concat, logits_vis, logits_syb

    def _two_branches(self, run_vis, run_syb, on_gpu):
        """Runs the two branch models and the heads.  The branches are independent until the heads (AttModel_x3.py:529-531): on a
        GPU the symbolic branch is forked onto a second stream so that its kernels (and, through autograd, their backward) overlap
        the visual branch's -- the decoders' launch-bound M = B kernels of one branch hide under the encoder GEMMs of the other.
        Inside a CUDA-graph capture this becomes two parallel branches of the graph."""
        if not (getattr(self, "concurrent_branches", True) and on_gpu):
            fea_vis_grid, fea_syb = run_vis(), run_syb()
            return self.answer_logits(Fn.bucket_mark(fea_vis_grid, (id(self), "heads")), Fn.bucket_mark(fea_syb, (id(self), "heads")))
        main = torch.cuda.current_stream()
        side = _SIDE_STREAMS.get(main.device)
        if side is None:
            side = _SIDE_STREAMS[main.device] = torch.cuda.Stream(device=main.device)
        if _BRANCH_SMS is not None:  # experiment: static split of the SMs between the two branches' persistent GEMMs
            for st, n in ((main, _BRANCH_SMS[0]), (side, _BRANCH_SMS[1])):
                ops.set_stream_sm_limit(st, n)
                ops.set_stream_sm_limit(Fn.wgrad_stream_of(st), min(n, Fn.SIDE_GEMM_SMS) if n else 0)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            fea_syb = run_syb()
        fea_vis_grid = run_vis()
        main.wait_stream(side)
        fea_syb.record_stream(main)
        return self.answer_logits(Fn.bucket_mark(fea_vis_grid, (id(self), "heads")), Fn.bucket_mark(fea_syb, (id(self), "heads")))

    def encoder_step(self, vis_fea, vis_mask, q_ipt, q_mask, q_graph, syb_ipt, macro_mask, macro_graph, decMask=True):
        """The two branch models + heads, with `syb_ipt` [B,M,2048] given (what MIL_NCE hands to att_syb at AttModel_x3.py:530)."""
        return self._two_branches(lambda: self.att_vis_grid(vis_fea, vis_mask, q_ipt, q_graph, q_mask, decMask),
                                  lambda: self.att_syb(syb_ipt, macro_mask, macro_graph, q_ipt, q_graph, q_mask, decMask), vis_fea.is_cuda)

    def forward_compact(self, c, decMask=True, mcb=False):
        """forward() on the loader's compact hand-off (collate.compact_batch): bf16 region features, per-sample lengths instead of
        the prefix-block masks, bit-packed adjacency.  Returns what forward() returns on the dense batch, bit for bit (the masks the
        kernels see are identical)."""
        if mcb == True:  # noqa: E712
            raise NotImplementedError("mcb=True needs torch.rfft (removed from PyTorch); the reference cannot run it either")
        vis_fea = c["vis_fea"]
        if vis_fea.dtype != torch.bfloat16:
            vis_fea = _CastBf16.apply(vis_fea)
        e = torch.empty((vis_fea.shape[0], 0), device=vis_fea.device)  # only_obj: main_itp_ddp_tar_super_node.py:290-308
        mil = {}

        def run_syb():
            # MIL_NCE feeds only the symbolic branch (AttModel_x3.py:525, 530): it runs at the head of that branch's stream, forward and
            # -- through autograd -- backward, next to the visual branch instead of in front of both
            new_macro_ipt, mil["obj"], mil["rel"] = self.MIL_NCE(vis_fea, c["macro_node_ipt"], c["macro_obj_loc_ipt"], c["micro_positive_obj_ipt"],
                                                                 c["micro_negative_obj_ipt"], c["micro_obj_mask"], e, e, e, e)
            return self.att_syb.forward_compact(new_macro_ipt, c["macro_len"], c["macro_graph_bits"], c["q_ipt"], c["q_graph_bits"], c["q_len"],
                                                decMask)
        logits = self._two_branches(
            lambda: self.att_vis_grid.forward_compact(vis_fea, c["vis_len"], c["q_ipt"], c["q_graph_bits"], c["q_len"], decMask),
            run_syb, vis_fea.is_cuda)
        if vis_fea.is_cuda and isinstance(mil["obj"], torch.Tensor):
            mil["obj"].record_stream(torch.cuda.current_stream())
        return (*logits, mil["obj"], mil["rel"])

    def forward(self, vis_fea, vis_mask, q_ipt, q_mask, q_graph, macro_ipt, macro_mask, macro_graph, macro_obj_loc,
                micro_positive_obj, micro_negative_obj, micro_obj_mask, micro_positive_rel, micro_negative_rel,
                micro_positive_rel_loc, micro_negative_rel_loc, decMask=True, mcb=False):
        if mcb == True:  # noqa: E712
            raise NotImplementedError("mcb=True needs torch.rfft (removed from PyTorch); the reference cannot run it either")
        if vis_fea.dim() == 4:  # bs x gridx x gridy x fea_size
            vis_fea = vis_fea.reshape(-1, vis_fea.size(1) * vis_fea.size(2), vis_fea.size(3))
        if vis_fea.dtype != torch.bfloat16:
            vis_fea = _CastBf16.apply(vis_fea)  # staged once: MIL_NCE's vis_mlp and the visual branch's syb_mlp2 both read it
        mil = {}

        def run_syb():  # MIL_NCE at the head of the symbolic branch's stream (it feeds only that branch, AttModel_x3.py:525, 530)
            new_macro_ipt, mil["obj"], mil["rel"] = self.MIL_NCE(vis_fea, macro_ipt, macro_obj_loc, micro_positive_obj, micro_negative_obj,
                                                                 micro_obj_mask, micro_positive_rel, micro_negative_rel,
                                                                 micro_positive_rel_loc, micro_negative_rel_loc)
            return self.att_syb(new_macro_ipt, macro_mask, macro_graph, q_ipt, q_graph, q_mask, decMask)
        logits_concat, logits_vis, logits_syb = self._two_branches(
            lambda: self.att_vis_grid(vis_fea, vis_mask, q_ipt, q_graph, q_mask, decMask), run_syb, vis_fea.is_cuda)
        if vis_fea.is_cuda and isinstance(mil["obj"], torch.Tensor):
            mil["obj"].record_stream(torch.cuda.current_stream())
        return logits_concat, logits_vis, logits_syb, mil["obj"], mil["rel"]


def answer_loss(logits_concat, logits_vis, logits_syb, answer, epsilon=0.1):
    """The train loop's loss (main_itp_ddp_tar_super_node.py:335-345) as one fused kernel with its gradient."""
    return Fn.AnswerLossFn.apply(logits_concat, logits_vis, logits_syb, answer, epsilon)
