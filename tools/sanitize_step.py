#!/usr/bin/env python
"""Small eager training steps through every kernel family of the path (bound trainer: fused decoder, MIL_NCE, compact masks,
tcgen05 GEMMs / attention, LayerNorm, deferred Adam) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck python tools/sanitize_step.py
(one tool per gpurun call, B200_PROFILING.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "structured-alignment-vqa_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from savqa_b200 import _lib, collate, synthetic, train  # noqa: E402

_lib.require_device()
big = len(sys.argv) > 1 and sys.argv[1] == "big"
cfg = dict(synthetic.GQA_SHAPED, ncls=64) if big else dict(synthetic.GQA_SHAPED, V=40, Q=8, M=60, ncls=64)  # T = 48 / 68: shared-tile attention backward
B = 8 if big else 4
model = synthetic.build_model(cfg, vocab_rows=2000).cuda()
c = {k: v.cuda() for k, v in collate.compact_batch(synthetic.make_batch(cfg, B, seed=3, vocab_rows=2000)).items()}
tr = train.EncoderTrainer(model, lr=1e-4, step="compact")
losses = [float(tr.step(c)) for _ in range(2)]
tr.flush_tables()
torch.cuda.synchronize()
print("sanitize_step: losses", losses, "engines", _lib.launch_counts())
assert all(l == l for l in losses)
