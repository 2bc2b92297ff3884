"""Batch-sharded inference and the evaluation loop of the graph-guided encoder path.

Replaces, for `--model_v 3`, the reference's `eval()` (models/main_itp_ddp_tar_super_node.py:60-142) -- the stale
`eval_itp_grid_ddp_tar_gt.py` calls a constructor / forward signature no AttModel in the reference has (SURVEY.md section 0,
fact 10).  Inference shards by batch: every rank runs its own samples, there is NO collective on the data path; only the three
per-epoch scalars (loss, correct, count) are gathered, exactly as the reference does (main...:382-392).

  * InferenceRunner: the forward pass captured once into a CUDA graph over static input buffers (the bf16 weight staging is
    cached outside the graph: weights do not change), with the same double-buffered pinned-host -> device input path as the
    trainer (stage / commit / replay).
  * evaluate(): loss (label-smoothed 3-head loss, + the MIL-NCE term when asked) and accuracy with the reference's counting rule
    (`pred[nonzero(answer)] == answer[nonzero(answer)]`: samples whose answer id is 0 are never counted as correct).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import AttModel_x3 as A

ENCODER_KEYS = ("vis_fea", "vis_fea_mask", "q_ipt", "q_ipt_mask", "q_ipt_graph", "syb_ipt", "macro_node_mask", "macro_graph_ipt")
FULL_KEYS = ("vis_fea", "vis_fea_mask", "q_ipt", "q_ipt_mask", "q_ipt_graph", "macro_node_ipt", "macro_node_mask", "macro_graph_ipt",
             "macro_obj_loc_ipt", "micro_positive_obj_ipt", "micro_negative_obj_ipt", "micro_obj_mask")
COMPACT_KEYS = ("vis_fea", "vis_len", "q_ipt", "q_len", "q_graph_bits", "macro_node_ipt", "macro_len", "macro_graph_bits",
                "macro_obj_loc_ipt", "micro_positive_obj_ipt", "micro_negative_obj_ipt", "micro_obj_mask")


def forward_batch(model: A.AttModel, b: Dict[str, torch.Tensor], dec_mask: bool = True, full: Optional[bool] = None):
    """(logits_concat, logits_vis, logits_syb, mil_nce_obj | None) of one collate_fn-shaped batch.  full=True runs the reference's
    16-argument AttModel.forward (MIL_NCE produces the symbolic node features, AttModel_x3.py:525-530); full=False takes
    `syb_ipt` [B,M,2048] as given (the encoder path alone)."""
    if "vis_len" in b:  # the loader's compact hand-off (collate.compact_batch)
        lc, lv, ls, mil_obj, _ = model.forward_compact(b, decMask=dec_mask)
        return lc, lv, ls, mil_obj
    if full is None:
        full = "micro_positive_obj_ipt" in b  # MIL_NCE's inputs are there: the reference's whole forward
    if full:
        e = b.get("micro_positive_rel_ipt")
        if e is None:
            e = torch.empty((b["vis_fea"].shape[0], 0), device=b["vis_fea"].device)  # only_obj: main...:290-308
        lc, lv, ls, mil_obj, _ = model(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["macro_node_ipt"],
                                       b["macro_node_mask"], b["macro_graph_ipt"], b["macro_obj_loc_ipt"], b["micro_positive_obj_ipt"],
                                       b["micro_negative_obj_ipt"], b["micro_obj_mask"], e, e, e, e, decMask=dec_mask, mcb=False)
        return lc, lv, ls, mil_obj
    lc, lv, ls = model.encoder_step(b["vis_fea"], b["vis_fea_mask"], b["q_ipt"], b["q_ipt_mask"], b["q_ipt_graph"], b["syb_ipt"],
                                    b["macro_node_mask"], b["macro_graph_ipt"], dec_mask)
    return lc, lv, ls, None


class InferenceRunner:
    def __init__(self, model: A.AttModel, dec_mask: bool = True, full: bool = False, keys: Optional[Sequence[str]] = None):
        self.model, self.dec_mask, self.full = model, dec_mask, full
        self.keys = tuple(keys) if keys is not None else (FULL_KEYS if full else ENCODER_KEYS)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static: Optional[Dict[str, torch.Tensor]] = None
        self.out: Optional[Tuple[torch.Tensor, ...]] = None
        self.staging = None
        self.launches_per_batch = 0

    def run(self, batch: Dict[str, torch.Tensor]):
        """One eager forward pass (device-resident inputs)."""
        self.model.eval()
        with torch.no_grad():
            return forward_batch(self.model, batch, self.dec_mask, self.full)

    def capture(self, batch: Dict[str, torch.Tensor], warmup: int = 2) -> None:
        from . import _lib
        self.static = {k: batch[k].clone() for k in self.keys}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                n0 = _lib.launch_count
                self.run(self.static)  # also stages (and caches) the bf16 weight packs: they stay out of the graph
                self.launches_per_batch = _lib.launch_count - n0
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self.run(self.static)

    def stage(self, batch: Dict[str, torch.Tensor]) -> None:
        """Host (pinned) batch -> staging buffers on a copy stream while the previous batch computes."""
        if self.staging is None:
            self.staging = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged = torch.cuda.Event()
            self._free = torch.cuda.Event()
            self._free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._free)
        with torch.cuda.stream(cs):
            for k in self.keys:
                self.staging[k].copy_(batch[k], non_blocking=True)
            self._staged.record(cs)

    def commit(self) -> None:
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for k in self.keys:
            self.static[k].copy_(self.staging[k], non_blocking=True)
        self._free.record(cur)

    def replay(self):
        self.graph.replay()
        return self.out


def evaluate(model: A.AttModel, batches: Iterable[Dict[str, torch.Tensor]], dec_mask: bool = True, with_milnce_loss: bool = False,
             process_group=None):
    """The reference's eval() for --model_v 3 (main_itp_ddp_tar_super_node.py:60-142 + the gather at :382-392): returns
    (mean loss over ranks, correct, count).  One device -> host read per CALL, not per batch (the reference syncs on
    `loss.cpu().item()` every batch)."""
    model.eval()
    dev = None
    loss_sum = n_batches = None
    correct = count = None
    with torch.no_grad():
        for b in batches:
            if dev is None:
                dev = b["vis_fea"].device
                loss_sum = torch.zeros((), device=dev)
                correct = torch.zeros((), device=dev)
                n_batches, count = 0, 0
            lc, lv, ls, mil_obj = forward_batch(model, b, dec_mask)
            loss = A.answer_loss(lc, lv, ls, b["answer"])
            if with_milnce_loss and mil_obj is not None:
                loss = loss - mil_obj  # mil_nce_loss = -mil_nce_obj (only_obj), main...:108-111, 130-131
            lsm = (torch.log_softmax(lv, -1) + torch.log_softmax(ls, -1) + torch.log_softmax(lc, -1)) / 3
            pred = lsm.argmax(1)
            nz = b["answer"] != 0  # `pred_lbls[torch.nonzero(answer)] == answer[torch.nonzero(answer)]`, main...:125
            correct += ((pred == b["answer"]) & nz).sum()
            bs = b["vis_fea"].shape[0]
            loss_sum += loss * bs   # AverageMeter.update(loss, batch_size)
            count += bs
            n_batches += 1
    if dev is None:
        return 0.0, 0, 0
    vals = torch.stack([loss_sum / max(count, 1), correct, torch.tensor(float(count), device=dev)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        got = [torch.zeros_like(vals) for _ in range(dist.get_world_size(process_group))]
        dist.all_gather(got, vals, group=process_group)
        st = torch.stack(got)
        return float(st[:, 0].mean()), int(st[:, 1].sum()), int(st[:, 2].sum())
    v = vals.tolist()
    return float(v[0]), int(v[1]), int(v[2])
