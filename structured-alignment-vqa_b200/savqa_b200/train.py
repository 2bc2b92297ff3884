"""Data-parallel training step of the graph-guided encoder path: one process per GPU.

Replaces, for this path, the reference's `DistributedDataParallel(find_unused_parameters=True)` + dense
`torch.optim.Adam` over 470 M parameters (main_itp_ddp_tar_super_node.py:203-206, 363-366):

  * the dense parameters that the step actually uses live in ONE flat fp32 buffer (parameter views), their gradients
    in another: one NCCL all-reduce (AVG) over NVLink and one fused Adam kernel per step;
  * the 407000 x 300 word tables exchange row-sparse gradients: all-gather of (row id, row gradient) lists, local
    scatter-add, and DEFERRED row-wise Adam (savqa_adam_rows): a row replays the zero-gradient updates it missed the next time it
    is read or updated, so the tables follow dense torch.optim.Adam (the reference's optimizer) without touching 1.5 GB of
    parameters, gradients and moments every step; flush_tables() brings every row up to date before a checkpoint.
    `rowsparse=False` keeps the tables in the dense flat buffers (dense gradients, dense Adam) for comparison;
  * the whole step (bf16 weight staging, forward, backward, all-reduce, optimizer) can be captured into one CUDA graph
    and replayed, which removes the Python / launch overhead of ~1500 kernel launches per step.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import AttModel_x3 as A
from . import functional as Fn
from . import ops

import os
import re
from collections import Counter


class GradReducer:
    """Bucketed gradient all-reduce that starts inside the backward pass (the role of DistributedDataParallel's reducer,
    main_itp_ddp_tar_super_node.py:203).  The flat gradient buffer is laid out bucket by bucket -- heads, each branch's
    decoder, each encoder block, the rest -- and the model reports, through functional.BucketMarkFn / MemoryJoinFn, the moment
    the last gradient kernel of a bucket has been LAUNCHED; the bucket's all-reduce is then enqueued on NCCL's stream behind
    the compute and weight-gradient streams that feed it, and overlaps the remaining backward.  finish() reduces whatever
    was not reported and hands back the (range, work) list for the optimizer to follow."""

    def __init__(self, flat_grad: torch.Tensor, buckets, need, group, world: int = 2, apply_fn=None, merged: bool = False):
        self.flat_grad, self.buckets, self.need, self.group = flat_grad, buckets, need, group
        self.world = world
        # merged (one GPU): ONE optimizer launch over all reported buckets, behind the last of them -- it then runs next to the input
        # stages' / MIL_NCE's backward and the word-table updates (a chain of small kernels on a nearly idle GPU) instead of after them
        self.merged = merged and apply_fn is not None and flat_grad.is_cuda
        # apply_fn(lo, hi): the optimizer update of flat range [lo, hi).  Under graph capture it is launched per bucket, right
        # behind the bucket's all-reduce (world 1: behind its last gradient kernel) on a stream of its own, so that the
        # HBM-bound Adam kernel runs under the rest of the backward pass instead of after it.  Safe: once a bucket's gradients
        # are final every kernel that reads the bucket's weights in this step (its dgrad GEMMs) has completed.
        self.apply_fn = apply_fn
        self.got, self.works, self.done, self.pending = Counter(), [], set(), []
        self.launch_stream = torch.cuda.Stream() if flat_grad.is_cuda else None
        self.apply_stream = torch.cuda.Stream() if (flat_grad.is_cuda and apply_fn is not None) else None

    def begin_step(self) -> None:
        self.got, self.works, self.done, self.pending = Counter(), [], set(), []

    @staticmethod
    def _stage(key) -> tuple:
        """Position of a bucket in the backward pass: heads, decoders, encoder blocks last to first."""
        kind = key[1]
        if kind == "heads":
            return (0, 0)
        if kind == "dec":
            return (1, 0)
        return (2, -int(key[2]))

    def ready(self, key) -> None:
        if key not in self.buckets or key in self.done:
            return
        self.got[key] += 1
        if self.got[key] < self.need.get(key, 1):
            return
        lo, hi = self.buckets[key]
        self.done.add(key)
        if hi <= lo:
            return
        if self.launch_stream is None:
            self.works.append((lo, hi, self._all_reduce(lo, hi), False))
            return
        cur = torch.cuda.current_stream()
        w = Fn._WGRAD_STREAMS.get((cur.device_index, cur.cuda_stream))
        # the heads run on streams of their own (AttModel.answer_logits): their bucket waits for every weight-gradient stream in use
        extra = [s_ for s_ in Fn._WGRAD_DIRTY if s_ is not w] if key[1] == "heads" else []
        if self.merged or torch.cuda.is_current_stream_capturing():
            # Inside a graph capture the HOST order of the launches is irrelevant to when the kernels run -- but collectives of
            # one communicator execute in launch order, and autograd walks the whole visual branch before the symbolic one: launched
            # from here, every symbolic-branch bucket would queue behind the visual branch's LAST bucket (seen in the 2-GPU kernel
            # trace: half of the all-reduce volume ran after the backward pass).  Record the readiness events now; finish() launches
            # the buckets in the order in which the GPU completes them (stage by stage, both branches alternating).
            evs = [torch.cuda.Event()]
            evs[0].record(cur)
            for s_ in ([w] if w is not None else []) + extra:
                evs.append(torch.cuda.Event())
                evs[-1].record(s_)
            self.pending.append((self._stage(key), len(self.pending), lo, hi, evs))
            return
        r = self.launch_stream
        r.wait_stream(cur)
        for s_ in ([w] if w is not None else []) + extra:
            r.wait_stream(s_)
        with torch.cuda.stream(r):
            work = self._all_reduce(lo, hi)
        self.works.append((lo, hi, work, False))

    def _all_reduce(self, lo: int, hi: int):
        if self.world <= 1:
            return None
        return dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True)

    def _launch_pending(self) -> bool:
        r, a = self.launch_stream, self.apply_stream
        if self.merged and self.pending:
            lo_all, hi_all = min(p_[2] for p_ in self.pending), max(p_[3] for p_ in self.pending)
            covered = sorted((p_[2], p_[3]) for p_ in self.pending)
            if all(covered[i][1] >= covered[i + 1][0] for i in range(len(covered) - 1)):  # the buckets tile one range of the buffer
                for p_ in self.pending:
                    for ev in p_[4]:
                        a.wait_event(ev)
                with torch.cuda.stream(a):
                    self.apply_fn(lo_all, hi_all)
                self.works.append((lo_all, hi_all, None, True))
                self.pending = []
                return False  # the launch stream took no part (under graph capture it must not be waited for then)
        for _, _, lo, hi, evs in sorted(self.pending, key=lambda t: (t[0], t[1])):
            for ev in evs:
                r.wait_event(ev)
            with torch.cuda.stream(r):
                work = self._all_reduce(lo, hi)
            applied = False
            if a is not None:
                if work is None:
                    for ev in evs:
                        a.wait_event(ev)
                with torch.cuda.stream(a):
                    if work is not None:
                        work.wait()
                    self.apply_fn(lo, hi)
                applied = True
            self.works.append((lo, hi, work, applied))
        self.pending = []
        return True

    def finish(self, chunks: int):
        """All-reduces the ranges no bucket report covered (in `chunks` pieces so that the optimizer can follow one piece
        behind) and returns every (lo, hi, work | None, already_applied) in launch order."""
        if self.launch_stream is not None:
            if self._launch_pending():
                torch.cuda.current_stream().wait_stream(self.launch_stream)
        covered = sorted(self.buckets[k] for k in self.done)
        gaps, pos, n = [], 0, self.flat_grad.numel()
        for lo, hi in covered:
            if lo > pos:
                gaps.append((pos, lo))
            pos = max(pos, hi)
        if pos < n:
            gaps.append((pos, n))
        total = sum(hi - lo for lo, hi in gaps)
        piece = max((total // max(chunks, 1) + 7) // 8 * 8, 8)
        for lo, hi in gaps:
            p = lo
            while p < hi:
                q = min(hi, p + piece)
                self.works.append((p, q, self._all_reduce(p, q), False))
                p = q
        return self.works


_BUCKET_RE = re.compile(r"^(att_vis_grid|att_syb)\.(enc|dec)_[a-z_]+_(\d+)\.")
#: encoder blocks per all-reduce bucket (1: one 11.5 MB bucket per block and branch; 2 / 3: fewer, larger collectives)
# 3 encoder blocks per bucket (35 MB of fp32 gradients): measured on 8 B200 against 1 / 2 blocks per bucket (8.27 vs 8.38 ms/step) --
# fewer, larger all-reduces spend less of the backward pass in NCCL's latency-bound regime (profiles/r2_bench_8gpu_variants.txt)
BUCKET_BLOCKS = max(1, int(os.environ.get("SAVQA_BUCKET_BLOCKS", "3")))


def bucket_key_of(model, name: str):
    """Which readiness bucket a parameter belongs to (see GradReducer): matches the marks placed in AttModel_x3._Branch."""
    m = _BUCKET_RE.match(name)
    if m:
        branch = getattr(model, m.group(1))
        return (id(branch), "enc", int(m.group(3)) // BUCKET_BLOCKS) if m.group(2) == "enc" else (id(branch), "dec")
    if name.split(".")[0] in ("cls", "cls_vis", "cls_syb"):
        return (id(model), "heads")
    return None


#: inputs of one step, by input format
STEP_KEYS = ("vis_fea", "vis_fea_mask", "q_ipt", "q_ipt_mask", "q_ipt_graph", "syb_ipt", "macro_node_mask", "macro_graph_ipt", "answer")
FULL_KEYS = ("vis_fea", "vis_fea_mask", "q_ipt", "q_ipt_mask", "q_ipt_graph", "macro_node_ipt", "macro_node_mask", "macro_graph_ipt",
             "macro_obj_loc_ipt", "micro_positive_obj_ipt", "micro_negative_obj_ipt", "micro_obj_mask", "answer")
from .collate import COMPACT_KEYS  # noqa: E402  (full step on the loader's compact hand-off)


class EncoderTrainer:
    def __init__(self, model: A.AttModel, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, rowsparse: bool = True,
                 process_group=None, dec_mask: bool = True, step: str = "encoder", with_milnce_loss: bool = True):
        """step: "encoder" -- the two branch models + heads with `syb_ipt` given (AttModel.encoder_step);
                 "full"    -- the train script's 16-argument AttModel.forward, MIL_NCE included, on collate_fn's dense batch
                              (main_itp_ddp_tar_super_node.py:321-325), loss += -mil_nce_obj when with_milnce_loss (:359-360);
                 "compact" -- the same full step on the loader's compact hand-off (collate.compact_batch)."""
        assert step in ("encoder", "full", "compact")
        self.step_kind, self.with_milnce_loss = step, with_milnce_loss
        self.keys = {"encoder": STEP_KEYS, "full": FULL_KEYS, "compact": COMPACT_KEYS}[step]
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.rowsparse = rowsparse
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        # The word tables' row lists travel on a communicator of their own (NCCL): collectives of ONE communicator run in launch
        # order, so on the gradient communicator a table's all-gathers would queue behind every bucket all-reduce that was launched
        # before them -- on this one they overlap the rest of the backward pass and the last buckets.  (A caller-supplied subgroup,
        # gloo, or SAVQA_TABLE_COMM=0: the gradient group is used.)
        self.table_pg = process_group
        if (self.world > 1 and process_group is None and dist.get_backend() == "nccl"
                and os.environ.get("SAVQA_TABLE_COMM", "1") != "0"):
            self.table_pg = dist.new_group(backend="nccl")
        self._table_streams: Dict[int, "torch.cuda.Stream"] = {}
        self._table_busy = []
        self.dec_mask = dec_mask
        self.step_count = 0
        self.allreduce_chunks = 6
        self.overlap_allreduce = os.environ.get("SAVQA_OVERLAP_ALLREDUCE", "1") != "0"  # world > 1: see GradReducer
        self._debug_skip_allreduce = os.environ.get("SAVQA_DEBUG_SKIP_ALLREDUCE", "0") == "1"  # timing experiments only
        self.tables = [model.att_vis_grid.syb_emb, model.att_syb.syb_emb] + ([model.MIL_NCE.syb_emb] if step != "encoder" else [])
        self.flat_param: Optional[torch.Tensor] = None
        self.row_state = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static: Optional[Dict[str, torch.Tensor]] = None
        self.static_loss: Optional[torch.Tensor] = None
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------
    def _forward_backward(self, b: Dict[str, torch.Tensor]) -> torch.Tensor:
        mil_obj = None
        if self.step_kind == "encoder":
            logits = self.model.encoder_step(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["syb_ipt"],
                                             b["macro_node_mask"], b["macro_graph_ipt"], self.dec_mask)
        elif self.step_kind == "full":
            e = torch.empty((b["vis_fea"].shape[0], 0), device=b["vis_fea"].device)  # only_obj: main...:290-308
            *logits, mil_obj, _ = self.model(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["macro_node_ipt"],
                                             b["macro_node_mask"], b["macro_graph_ipt"], b["macro_obj_loc_ipt"], b["micro_positive_obj_ipt"],
                                             b["micro_negative_obj_ipt"], b["micro_obj_mask"], e, e, e, e, decMask=self.dec_mask, mcb=False)
        else:
            *logits, mil_obj, _ = self.model.forward_compact(b, decMask=self.dec_mask)
        loss = A.answer_loss(*logits, b["answer"])
        if mil_obj is not None and self.with_milnce_loss:
            loss = loss - mil_obj  # loss += mil_nce_loss, mil_nce_loss = -mil_nce_obj (main...:326-329, 359-360)
        self.last_mil_obj = mil_obj.detach() if mil_obj is not None else None
        zs = getattr(self, "_zero_done", None)
        if zs is not None:  # the flat gradient buffer was being zeroed next to the forward pass (see _step_impl)
            torch.cuda.current_stream().wait_event(zs)
            self._zero_done = None
        loss.backward()
        return loss.detach()

    def prepare(self, batch: Dict[str, torch.Tensor]) -> None:
        """Discovers (one dry run) which parameters the step gives gradients to, then moves them and their gradients
        into flat buffers.  The reference needs find_unused_parameters=True for the same reason."""
        model = self.model
        dev = batch["vis_fea"].device
        for t in self.tables:
            t._savqa_rowlog = Fn.RowGradLog() if self.rowsparse else None
        model.zero_grad(set_to_none=True)
        self._forward_backward(batch)
        for t in self.tables:
            if t._savqa_rowlog is not None:
                t._savqa_rowlog.clear()
        table_ids = {id(t.weight) for t in self.tables}
        used = [p for p in model.parameters() if p.grad is not None and id(p) not in table_ids]
        if not self.rowsparse:
            used += [t.weight for t in self.tables]
        used_ids = {id(p) for p in used}
        # layout: the groups the modules ask for ([Wq; Wk; Wv] of one attention as one [3C, C] block, ...) first, then the rest;
        # every group starts on a multiple of 8 elements (16-byte aligned rows in the bf16 mirror for TMA)
        order, placed = [], set()
        for mod in model.modules():
            groups = mod._savqa_groups() if hasattr(mod, "_savqa_groups") else []
            for g in groups:
                ids = [id(p) for p in g]
                if all(i in used_ids for i in ids) and not any(i in placed for i in ids):
                    if len(g) > 1 and any(p.numel() % 8 for p in g[:-1]):
                        continue  # members would lose their alignment; leave them separate
                    order.append(list(g))
                    placed.update(ids)
        order += [[p] for p in used if id(p) not in placed]
        # bucket by readiness in the backward pass (heads, decoders, encoder blocks last-to-first, the rest), so that each
        # bucket is one contiguous range of the flat gradient buffer that can be all-reduced as soon as it is final
        names = {id(p): k for k, p in model.named_parameters()}
        keys = []
        for g in order:
            k = bucket_key_of(model, names.get(id(g[0]), ""))
            if k is not None and k not in keys:
                keys.append(k)
        rank = {k: i for i, k in enumerate(keys)}
        order.sort(key=lambda g: rank.get(bucket_key_of(model, names.get(id(g[0]), "")), len(rank)))
        offsets, n = {}, 0
        self.bucket_ranges = {}
        for g in order:
            n = (n + 7) // 8 * 8
            k = bucket_key_of(model, names.get(id(g[0]), ""))
            for p in g:
                offsets[id(p)] = (n, p.numel())
                n += p.numel()
            if k is not None:
                lo, _ = self.bucket_ranges.get(k, (offsets[id(g[0])][0], 0))
                self.bucket_ranges[k] = (lo, (n + 7) // 8 * 8)
        n_pad = (n + 7) // 8 * 8
        self.dense: List[torch.nn.Parameter] = [p for g in order for p in g]
        self.flat_param = torch.zeros(n_pad, device=dev)
        self.flat_grad = torch.zeros(n_pad, device=dev)
        self.exp_avg = torch.zeros(n_pad, device=dev)
        self.exp_avg_sq = torch.zeros(n_pad, device=dev)
        self.flat_bf16 = torch.zeros(n_pad, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            for p in self.dense:
                o, k = offsets[id(p)]
                self.flat_param[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat_param[o:o + k].view(p.shape)
                p.grad = self.flat_grad[o:o + k].view(p.shape)
        if self.world > 1:
            # what DistributedDataParallel's constructor does (main_itp_ddp_tar_super_node.py:203): every rank starts from rank 0's
            # parameters -- the reference does not seed the ranks identically, and Xavier / Kaiming initialisation depends on the RNG
            dist.broadcast(self.flat_param, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
            for t in self.tables:
                dist.broadcast(t.weight.data, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        self.sync_mirror()
        # the modules' weight packs / LayerNorm sinks now read the mirror and write their gradients straight into flat_grad
        self.views = Fn.FlatViews(self.flat_param, self.flat_grad, self.flat_bf16, offsets)
        for mod in model.modules():
            if hasattr(mod, "_savqa_bind"):
                mod._savqa_bind(self.views)
        self.n_dense = n
        if self.rowsparse:
            self.row_state = []
            for t in self.tables:
                w = t.weight
                w.grad = None
                self.row_state.append(dict(grad=torch.zeros(w.shape, dtype=torch.int64, device=w.device), m=torch.zeros_like(w), v=torch.zeros_like(w),
                                           stamp=torch.zeros(w.shape[0], dtype=torch.int32, device=dev)))
        # {lr / (1 - b1^s), sqrt(1 - b2^s), s}: the step counter lives on the DEVICE and is advanced by a kernel at the head of
        # every step (ops.adam_advance, part of the captured graph), so a host that queues steps ahead of the GPU cannot hand a
        # step the scalars -- or the row stamp of savqa_adam_rows -- of another one
        self.dyn = torch.zeros(3, device=dev)
        if self.rowsparse:
            import functools
            for i, t in enumerate(self.tables):
                # deferred Adam, part 1: rows are caught up right before the gather that reads them, on that gather's stream
                t._savqa_rowlog.before_read = functools.partial(self._catch_up_ids, i)
                # part 2: a table's exchange (several ranks: all-gather of its row lists) and update start as soon as its gradient rows
                # exist, on a helper stream, under the rest of the backward pass
                eager = self.world == 1 or (dev.type == "cuda" and os.environ.get("SAVQA_TABLES_EAGER", "1") != "0")
                t._savqa_rowlog.on_grad = functools.partial(self._apply_table_now, i) if eager else None
        Fn.WEIGHT_EPOCH += 1
        self.reducer = None
        # Per-bucket Adam under the backward pass (GradReducer.apply_fn) is implemented but OFF: measured on B200 it moves the
        # 0.44 ms of Adam under the backward and the backward slows down by as much (1 GPU: 7.65 -> 7.65 ms; 2 GPUs: 8.07 ->
        # 8.39 ms) -- the step is bound by the aggregate HBM / SM time of its kernels, not by its critical path.
        per_bucket_adam = os.environ.get("SAVQA_ADAM_PER_BUCKET", "0") == "1" and self.flat_grad.is_cuda
        # One GPU: the Adam of everything but the input stages / MIL_NCE (96 % of the parameters) is launched ONCE, behind the last
        # encoder block's gradients -- next to the chain of small kernels that ends the backward pass, not after it
        early_adam = self.world == 1 and self.flat_grad.is_cuda and os.environ.get("SAVQA_EARLY_ADAM", "1") != "0" and not per_bucket_adam
        if (self.world > 1 or per_bucket_adam or early_adam) and self.overlap_allreduce and not self._debug_skip_allreduce:
            # both decoder outputs feed the heads; a bucket of several encoder blocks is final when the backward pass has left its
            # lowest block
            nb = max((getattr(b_, "num_blocks", 0) for b_ in (model.att_vis_grid, model.att_syb)), default=0)
            need = {k: (2 if k[1] == "heads" else (min(BUCKET_BLOCKS, nb - k[2] * BUCKET_BLOCKS) if k[1] == "enc" else 1))
                    for k in self.bucket_ranges}
            self.reducer = GradReducer(self.flat_grad, dict(self.bucket_ranges), need, self.pg, self.world,
                                       self._adam_range if (per_bucket_adam or early_adam) else None, merged=early_adam)

    def sync_mirror(self) -> None:
        """Re-derives the bf16 mirror from the fp32 parameters (after prepare(), or after load_state_dict wrote into them)."""
        ops.cast_bf16(self.flat_param.view(-1, 8), out=self.flat_bf16.view(-1, 8))

    def release(self) -> None:
        """Detaches the modules from the flat buffers (they go back to per-call weight staging and autograd gradients)."""
        for mod in self.model.modules():
            if hasattr(mod, "_savqa_bind"):
                mod._savqa_bind(None)

    # ------------------------------------------------------------------------------------------------
    def _exchange_and_apply(self) -> None:
        b1, b2 = self.betas
        if self.flat_param.is_cuda:
            Fn.join_wgrad_streams()  # weight-gradient GEMMs run on side streams during the backward pass
        step = max(self.step_count, 1)
        works = None
        if self.world > 1 and not self._debug_skip_allreduce:
            # the all-reduce travels in pieces on NCCL's stream -- the buckets the backward pass reported (GradReducer), then the
            # remainder -- while the fused Adam kernel follows one piece behind on the compute stream.  Launched BEFORE the word
            # tables' all-gathers: collectives of one communicator run in launch order, and those wait for the end of the backward
            if self.reducer is not None:
                works = self.reducer.finish(2)  # what no bucket covers is small (input MLPs, tables): two pieces, latency bound
            else:
                n = self.flat_grad.numel()
                k = self.allreduce_chunks
                bounds = [(n * i // k) // 8 * 8 for i in range(k)] + [n]
                works = [(lo, hi, dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.AVG, group=self.pg, async_op=True), False)
                         for lo, hi in zip(bounds[:-1], bounds[1:])]
        elif self.reducer is not None:
            works = self.reducer.finish(1)  # single GPU: per-bucket Adam launched under the backward pass, the remainder here
        rows_done = False
        if self.rowsparse and self.flat_param.is_cuda:
            # the word tables' row updates (and, multi-GPU, the all-gathers of their row lists) touch nothing the flat gradient
            # all-reduce / Adam kernel touch -> on a helper stream, under them
            cur = torch.cuda.current_stream()
            aside = Fn.wgrad_stream_of(cur)
            aside.wait_stream(cur)
            with torch.cuda.stream(aside):
                self._apply_rows()
            rows_done, self._rows_stream = True, aside
        if works is not None:
            for lo, hi, w, applied in works:
                if applied:
                    continue
                if w is not None:
                    w.wait()
                self._adam_range(lo, hi)
            if self.reducer is not None and self.reducer.apply_stream is not None:
                torch.cuda.current_stream().wait_stream(self.reducer.apply_stream)
        else:
            ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, b1, b2, self.eps, step,
                          dyn=self.dyn, param_bf16=self.flat_bf16)
        if self.rowsparse and not rows_done:
            self._apply_rows()
        elif rows_done:
            torch.cuda.current_stream().wait_stream(self._rows_stream)
        for s_ in self._table_busy:  # tables whose exchange + update started inside the backward pass (several ranks)
            torch.cuda.current_stream().wait_stream(s_)
        self._table_busy = []

    def _adam_range(self, lo: int, hi: int) -> None:
        b1, b2 = self.betas
        ops.adam_step(self.flat_param[lo:hi], self.flat_grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], self.lr, b1, b2,
                      self.eps, max(self.step_count, 1), dyn=self.dyn, param_bf16=self.flat_bf16[lo:hi])

    def _catch_up_ids(self, i: int, ids: torch.Tensor) -> None:
        """Deferred Adam, part 1 (RowGradLog.before_read): the rows a gather is about to read replay the zero-gradient updates dense
        Adam applied to them while they were absent from the batches (savqa_adam_rows, apply=0)."""
        b1, b2 = self.betas
        st = self.row_state[i]
        ops.adam_rows(self.tables[i].weight.data, None, st["m"], st["v"], st["stamp"], ids.reshape(-1), self.lr, b1, b2, self.eps,
                      max(self.step_count, 1), dyn=self.dyn, apply=False)

    def _apply_table(self, i: int) -> None:
        """Deferred Adam, part 2 for one table: its (row id, row gradient) lists -> (all-gather over the ranks) -> scatter-add -> this
        step's update of the touched rows (rows that only another rank read are caught up here first)."""
        b1, b2 = self.betas
        t, st = self.tables[i], self.row_state[i]
        log = t._savqa_rowlog
        touched = []
        for idx, rows, scale, skip in log.pending:
            if self.world > 1:
                idx_all = torch.empty(self.world * idx.numel(), dtype=idx.dtype, device=idx.device)
                rows_all = torch.empty(self.world * rows.shape[0], rows.shape[1], dtype=rows.dtype, device=rows.device)
                dist.all_gather_into_tensor(idx_all, idx, group=self.table_pg)
                dist.all_gather_into_tensor(rows_all, rows, group=self.table_pg)
                idx, rows, scale = idx_all, rows_all, scale / self.world
            ops.scatter_add_rows(st["grad"], idx, rows, scale=scale, skip_row=skip)
            touched.append(idx.reshape(-1))
        if touched:  # ONE update per table and step, after every gradient list has been accumulated (a row is claimed once per step)
            ops.adam_rows(t.weight.data, st["grad"], st["m"], st["v"], st["stamp"], touched[0] if len(touched) == 1 else torch.cat(touched),
                          self.lr, b1, b2, self.eps, max(self.step_count, 1), dyn=self.dyn, apply=True)
        log.clear()

    def _apply_table_now(self, i: int) -> None:
        """RowGradLog.on_grad (one GPU): launch table i's update on the helper stream of the stream its gradient was produced on."""
        if self.row_state is None:
            return
        if not self.flat_param.is_cuda:
            self._apply_table(i)
            return
        cur = torch.cuda.current_stream()
        if self.world == 1:
            aside = Fn.wgrad_stream_of(cur)
            aside.wait_stream(cur)
            if aside not in Fn._WGRAD_DIRTY:
                Fn._WGRAD_DIRTY.append(aside)  # joined with the weight-gradient streams before the optimizer (join_wgrad_streams)
        else:
            # a stream per table: the exchange of one table must not queue behind the update of another, and the optimizer's
            # all-reduces must not wait for it either (joined at the very end of the step, _exchange_and_apply)
            aside = self._table_streams.get(i)
            if aside is None:
                aside = self._table_streams[i] = torch.cuda.Stream()
            aside.wait_stream(cur)
            self._table_busy.append(aside)
        for idx, rows, _, _ in self.tables[i]._savqa_rowlog.pending:
            idx.record_stream(aside)
            rows.record_stream(aside)
        with torch.cuda.stream(aside):
            self._apply_table(i)

    def flush_tables(self) -> None:
        """Brings EVERY row of the word tables up to date with the current step (before state_dict() / evaluation): after it the
        tables hold exactly what dense torch.optim.Adam would (main_itp_ddp_tar_super_node.py:206, 366, 428)."""
        if not self.rowsparse or self.flat_param is None:
            return
        b1, b2 = self.betas
        for t, st in zip(self.tables, self.row_state):
            ops.adam_rows(t.weight.data, None, st["m"], st["v"], st["stamp"], None, self.lr, b1, b2, self.eps, self.step_count + 1,
                          dyn=None, apply=False)

    def _apply_rows(self) -> None:
        """Updates of the tables whose gradient lists are still pending (every table when there are several ranks)."""
        for i in range(len(self.tables)):
            if self.tables[i]._savqa_rowlog.pending:
                self._apply_table(i)

    def _step_impl(self, b: Dict[str, torch.Tensor]) -> torch.Tensor:
        ops.adam_advance(self.dyn, self.lr, self.betas[0], self.betas[1])  # step += 1 on the device; read by the Adam kernels below
        if self.flat_grad.is_cuda:
            # 356 MB of zeros (48 us at HBM speed) that nothing reads before the backward pass: on a helper stream, next to the forward
            cur = torch.cuda.current_stream()
            if getattr(self, "_zero_stream", None) is None:
                self._zero_stream = torch.cuda.Stream()
            self._zero_stream.wait_stream(cur)
            with torch.cuda.stream(self._zero_stream):
                ops.fill_zero(self.flat_grad, max_blocks=96)  # a background trickle: does not hold the block scheduler (see the kernel)
                self._zero_done = torch.cuda.Event()
                self._zero_done.record(self._zero_stream)
        else:
            self.flat_grad.zero_()
        if self.reducer is not None:
            self.reducer.begin_step()
            Fn.GRAD_REDUCER = self.reducer
        try:
            loss = self._forward_backward(b)
        finally:
            Fn.GRAD_REDUCER = None
        self._exchange_and_apply()
        return loss

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One eager training step on device-resident inputs; returns the (device) loss."""
        if self.flat_param is None:
            self.prepare(batch)
        self.step_count += 1
        from . import _lib
        n0 = _lib.launch_count
        loss = self._step_impl(batch)
        self.launches_per_step = _lib.launch_count - n0
        Fn.WEIGHT_EPOCH += 1
        return loss

    # ------------------------------------------------------------------------------------------------
    def capture(self, batch: Dict[str, torch.Tensor], warmup: int = 3) -> None:
        """Captures the whole step into a CUDA graph over static input buffers."""
        if self.flat_param is None:
            self.prepare(batch)
        self.static = {k: batch[k].clone() for k in self.keys}
        Fn.FORCE_RESTAGE = True  # the bf16 weight staging must be part of the replayed graph
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):  # records only: the step counter moves when the graph is replayed
            self.static_loss = self._step_impl(self.static)
        Fn.WEIGHT_EPOCH += 1

    def load_static(self, batch: Dict[str, torch.Tensor], stream: Optional[torch.cuda.Stream] = None) -> None:
        """Copies a (pinned host or device) batch into the graph's static input buffers."""
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
        with ctx:
            for k in self.keys:
                self.static[k].copy_(batch[k], non_blocking=True)

    # ---- double-buffered host -> device input path (the loader hand-off of SURVEY.md 8(f2)) ---------------------
    def stage(self, batch: Dict[str, torch.Tensor]) -> None:
        """Starts the host -> device copy of the NEXT step's (pinned) batch on a copy stream, into a staging set of buffers,
        while the current step computes; commit() then moves it into the graph's static inputs with device-to-device copies
        (~0.1 ms for 164 MB) on the compute stream."""
        if getattr(self, "staging", None) is None:
            self.staging = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._staging_free)  # the previous commit() has consumed the staging buffers
        with torch.cuda.stream(cs):
            for k in self.keys:
                self.staging[k].copy_(batch[k], non_blocking=True)
            self._staged.record(cs)

    def commit(self) -> None:
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for k in self.keys:
            self.static[k].copy_(self.staging[k], non_blocking=True)
        self._staging_free.record(cur)

    def replay(self) -> torch.Tensor:
        self.step_count += 1  # host-side bookkeeping only: the kernels read the device counter
        self.graph.replay()
        return self.static_loss
