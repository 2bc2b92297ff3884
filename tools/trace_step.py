#!/usr/bin/env python
"""Kernel timeline of the graph-replayed training step (torch.profiler / CUPTI activity records): which stream runs what, when.
Writes gpurun_out/trace_step.json.gz (compact: name, stream, start, duration of every kernel of ONE replay) and prints the
per-stream busy time and the union busy time of the step.  usage: python tools/trace_step.py [--batch 128]"""
import argparse
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from savqa_b200 import _lib, synthetic, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
_lib.require_device()
dev = torch.device("cuda", local)
if world > 1:  # under torchrun: every rank runs the step, rank 0 records its own timeline (NCCL kernels included)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cfg = synthetic.GQA_SHAPED
model = synthetic.build_model(cfg, seed=0).to(dev)
model.train()
from savqa_b200 import collate  # noqa: E402
b = collate.compact_batch(synthetic.make_batch(cfg, args.batch, seed=100 + rank * 16))
dev_batch = {k: b[k].to(dev) for k in train.COMPACT_KEYS}
trainer = train.EncoderTrainer(model, lr=1e-4, rowsparse=True, step="compact")
trainer.prepare(dev_batch)
trainer.capture(dev_batch, warmup=2)
for _ in range(5):
    trainer.replay()
torch.cuda.synchronize()
import contextlib  # noqa: E402
with (profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) if rank == 0 else contextlib.nullcontext()) as prof:
    for _ in range(3):
        trainer.replay()
    torch.cuda.synchronize()
if rank != 0:
    torch.distributed.barrier()
    os._exit(0)
path = os.path.join(ROOT, "gpurun_out", "trace_full.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e and "dur" in e]
ks.sort(key=lambda e: e["ts"])
print("gpu activities:", len(ks))
if ks:
    # split into replays at the kernel that advances the optimizer's device-resident step counter (savqa_adam_advance: the
    # first activity of every replay)
    starts = [i for i, e in enumerate(ks) if "adam_advance_kernel" in e["name"]]
    lo = starts[1] if len(starts) >= 3 else 0
    hi = starts[2] if len(starts) >= 3 else len(ks)
    one = ks[lo:hi]
    t0 = one[0]["ts"]
    rows = [dict(n=e["name"].replace("void ", "").replace("savqa::(anonymous namespace)::", "")[:90], s=e["args"].get("stream", -1), t=round(e["ts"] - t0, 2), d=round(e["dur"], 2)) for e in one]
    with gzip.open(os.path.join(ROOT, "gpurun_out", "trace_step.json.gz"), "wt") as f:
        json.dump(rows, f)
    span = one[-1]["ts"] + one[-1]["dur"] - t0
    print(f"one replay: {len(one)} activities, span {span / 1e3:.3f} ms")
    by = {}
    for r in rows:
        by.setdefault(r["s"], []).append(r)
    for s_, rs in sorted(by.items()):
        print(f"  stream {s_}: {len(rs)} kernels, busy {sum(r['d'] for r in rs) / 1e3:.3f} ms")
    # union busy
    iv = sorted((r["t"], r["t"] + r["d"]) for r in rows)
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for a, b_ in iv[1:]:
        if a > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = a, b_
        else:
            cur_e = max(cur_e, b_)
    busy += cur_e - cur_s
    print(f"  union busy {busy / 1e3:.3f} ms, idle {(span - busy) / 1e3:.3f} ms")
os.remove(path)
if world > 1:
    torch.distributed.barrier()
    sys.stdout.flush()
    os._exit(0)
