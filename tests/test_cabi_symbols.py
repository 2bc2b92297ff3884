"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/savqa_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "savqa_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    from savqa_b200 import _lib
    return _lib.LIB_PATH


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(savqa_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for need in ("savqa_build_masks", "savqa_gather_rows", "savqa_scatter_add_rows", "savqa_gemm_bf16", "savqa_graph_attn_fwd",
                 "savqa_graph_attn_bwd", "savqa_residual_layernorm_fwd", "savqa_layernorm_bwd", "savqa_last_error"):
        assert need in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/savqa_b200.h but not exported by {lib_path}"
    lib.savqa_abi_version.restype = ctypes.c_int
    assert lib.savqa_abi_version() == 5


def test_python_binding_covers_the_header(lib_path):
    from savqa_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_struct_layouts_match_header():
    """ctypes mirrors of the two argument structs have the C sizes (LP64: 8-byte pointers / int64, 4-byte int / float)."""
    from savqa_b200 import _lib
    assert ctypes.sizeof(_lib.GemmEpilogue) == 16 + 12 * 8
    assert ctypes.sizeof(_lib.AttnArgs) == 11 * 8 + 8 * 4 + 19 * 8 + 8  # + scale_d (int, padded to 8)
