"""Loader -> device hand-off in COMPACT form.

The reference's collate_fn (models/data_loader_itp_bbox_super_node_onlyobj.py:341-445) emits dense int32 planes -- `vis_fea_mask`
[B,V,V], `macro_node_mask` / `macro_graph_ipt` [B,M,M], `q_ipt_mask` / `q_ipt_graph` [B,Q,Q] -- and fp32 region features, and the
train loop copies them to the device one by one (main_itp_ddp_tar_super_node.py:271-316).  Everything in those planes is either a
PREFIX BLOCK (ones on [:n,:n]: the three masks, :355-358, 369-372, 412-414) or a 0/1 adjacency matrix (:374-380, 416-418), so the
same information is a length per sample and one bit per edge slot:

    vis_fea            bf16 [B,V,2048]   (the first kernel that reads the features casts them to bf16 anyway)
    vis_len / macro_len / q_len   int32 [B]
    macro_graph_bits   int32 [B,M,ceil(M/32)],  q_graph_bits int32 [B,Q,ceil(Q/32)]   (bit j of word w of row r = edge r -> 32 w + j)
    word ids, object locations, micro_obj_mask, answer: unchanged (small)

~19.5 MB per 128-sample GQA-shaped batch instead of 164 MB (50.8 MB with MIL_NCE on the device); `savqa_build_masks_compact`
rebuilds the fp32 / bit-packed graphs on the device, bit for bit equal to the dense path.  compact_batch() is what a
collate_fn replacement calls (host tensors in, host tensors out); expand_batch() is its inverse (tests).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

COMPACT_KEYS = ("vis_fea", "vis_len", "q_ipt", "q_len", "q_graph_bits", "macro_node_ipt", "macro_len", "macro_graph_bits",
                "macro_obj_loc_ipt", "micro_positive_obj_ipt", "micro_negative_obj_ipt", "micro_obj_mask", "answer")


def pack_adjacency(adj: torch.Tensor) -> torch.Tensor:
    """0/1 [B,N,N] (any integer / bool dtype, host) -> int32 [B,N,ceil(N/32)], bit j of word w = adj[b, r, 32 w + j]."""
    a = (adj.cpu().numpy() != 0)
    B, N, N2 = a.shape
    w = (N2 + 31) // 32
    padded = np.zeros((B, N, w * 32), dtype=np.uint8)
    padded[:, :, :N2] = a
    packed = np.packbits(padded.reshape(B, N, w, 32), axis=-1, bitorder="little")  # [B,N,w,4] bytes, little endian
    return torch.from_numpy(packed.view("<u4").reshape(B, N, w).astype(np.int32))


def unpack_adjacency(bits: torch.Tensor, n: int) -> torch.Tensor:
    b = bits.cpu().numpy().astype(np.int32).view(np.uint32)
    B, N, w = b.shape
    by = b.reshape(B, N, w, 1).view(np.uint8)
    return torch.from_numpy(np.unpackbits(by, axis=-1, bitorder="little").reshape(B, N, w * 32)[:, :, :n].astype(np.int32))


def prefix_lengths(mask: torch.Tensor) -> torch.Tensor:
    """[B,N,N] prefix-block mask -> int32 [B]; raises if a mask is not ones on [:n,:n] and zero elsewhere (what collate_fn emits)."""
    m = (mask != 0)
    n = m[:, 0, :].sum(-1).to(torch.int32)
    ar = torch.arange(mask.shape[1])
    valid = ar[None, :] < n[:, None]
    if not torch.equal(m, valid[:, :, None] & valid[:, None, :]):
        raise ValueError("savqa_b200.collate: mask is not a prefix block (ones on [:n,:n]); the compact hand-off cannot carry it")
    return n


def compact_batch(b: Dict[str, torch.Tensor], pin: bool = False) -> Dict[str, torch.Tensor]:
    out = {
        "vis_fea": b["vis_fea"].to(torch.bfloat16),
        "vis_len": prefix_lengths(b["vis_fea_mask"]),
        "q_ipt": b["q_ipt"],
        "q_len": prefix_lengths(b["q_ipt_mask"]),
        "q_graph_bits": pack_adjacency(b["q_ipt_graph"]),
        "macro_node_ipt": b["macro_node_ipt"],
        "macro_len": prefix_lengths(b["macro_node_mask"]),
        "macro_graph_bits": pack_adjacency(b["macro_graph_ipt"]),
    }
    for k in ("macro_obj_loc_ipt", "micro_positive_obj_ipt", "micro_negative_obj_ipt", "micro_obj_mask", "answer", "syb_ipt"):
        if k in b:
            out[k] = b[k]
    if pin:
        out = {k: v.contiguous().pin_memory() for k, v in out.items()}
    return out


def expand_batch(c: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of compact_batch (features come back as fp32 images of their bf16 values)."""
    V, Q, M = c["vis_fea"].shape[1], c["q_ipt"].shape[1], c["macro_node_ipt"].shape[1]

    def block(n, N):
        valid = torch.arange(N)[None, :] < n.cpu()[:, None].long()
        return (valid[:, :, None] & valid[:, None, :]).to(torch.int32)

    out = {k: v for k, v in c.items() if k not in ("vis_fea", "vis_len", "q_len", "q_graph_bits", "macro_len", "macro_graph_bits")}
    out.update(vis_fea=c["vis_fea"].float(), vis_fea_mask=block(c["vis_len"], V), q_ipt_mask=block(c["q_len"], Q),
               q_ipt_graph=unpack_adjacency(c["q_graph_bits"], Q), macro_node_mask=block(c["macro_len"], M),
               macro_graph_ipt=unpack_adjacency(c["macro_graph_bits"], M))
    return out
