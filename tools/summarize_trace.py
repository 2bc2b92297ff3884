#!/usr/bin/env python
"""Text summary of a step timeline written by tools/trace_step.py (gpurun_out/trace_step.json.gz): the phases of the step,
the per-kernel in-step durations (elapsed while sharing the GPU with the other streams, not exclusive time) and the
collectives.  usage: python tools/summarize_trace.py trace_step.json.gz"""
import collections
import gzip
import json
import re
import sys

rows = json.load(gzip.open(sys.argv[1], "rt"))
span = max(r["t"] + r["d"] for r in rows)
print(f"one replay of the captured step: {len(rows)} GPU activities, span {span / 1e3:.3f} ms, "
      f"{len(set(r['s'] for r in rows))} streams")


def window(sub):
    ts = [(r["t"], r["t"] + r["d"]) for r in rows if sub in r["n"]]
    return (min(t[0] for t in ts), max(t[1] for t in ts), len(ts)) if ts else None


print("\nphase markers (first start .. last end, us):")
for name, sub in (("encoder attention forward", "attn_fwd_tc"), ("decoder cross-attention forward", "attn_row1_fwd"),
                  ("loss", "answer_loss"), ("decoder cross-attention backward", "attn_row1_bwd"),
                  ("encoder attention backward", "attn_bwd_tc"), ("Adam", "adam_kernel"), ("NCCL", "nccl")):
    w = window(sub)
    if w:
        print(f"  {name:36s} {w[0]:8.0f} .. {w[1]:8.0f}   ({w[2]} launches)")
agg = collections.defaultdict(list)
for r in rows:
    m = re.match(r"([A-Za-z0-9_:]+(<[^>]*>)?)", r["n"])
    agg[m.group(1) if m else r["n"][:40]].append(r["d"])
print("\nkernel                                              n   total us   min    med    max   (in-step elapsed)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:30]:
    v2 = sorted(v)
    print(f"{k[:50]:50s} {len(v):3d} {sum(v):9.1f} {v2[0]:6.1f} {v2[len(v2) // 2]:6.1f} {v2[-1]:6.1f}")
print(f"sum of in-step durations {sum(sum(v) for v in agg.values()) / 1e3:.2f} ms over a span of {span / 1e3:.2f} ms")
