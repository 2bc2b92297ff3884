"""savqa_b200 -- B200-native (sm_100a) graph-guided attention encoder path of SA-VQA behind the reference's own
module API.  `savqa_b200.modules` mirrors reference `models/modules.py`; `savqa_b200.AttModel_x3` mirrors
`models/AttModel_x3.py`.  The compute lives in lib/libsavqa_b200.so (C ABI: include/savqa_b200.h)."""
import os as _os

# Several ranks: the captured step holds ~40 concurrent branches (two branch models, weight-gradient streams, per-bucket all-reduces,
# the word tables' exchanges).  With the driver's default of 8 hardware work queues independent branches share a queue and a
# collective that waits for its gradients holds up every node queued behind it (measured on 2 and 8 B200: the bucket all-reduces
# stalled behind the first table exchange until the end of the backward pass, profiles/r2_step_trace_8gpu_before.txt).  Read at
# context creation, so it has to be in the environment before the first CUDA call; a user's own setting wins.
if int(_os.environ.get("WORLD_SIZE", "1") or 1) > 1:
    _os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402

__version__ = "0.1.0"
