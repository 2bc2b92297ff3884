"""savqa_b200 -- B200-native (sm_100a) graph-guided attention encoder path of SA-VQA behind the reference's own
module API.  `savqa_b200.modules` mirrors reference `models/modules.py`; `savqa_b200.AttModel_x3` mirrors
`models/AttModel_x3.py`.  The compute lives in lib/libsavqa_b200.so (C ABI: include/savqa_b200.h)."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
