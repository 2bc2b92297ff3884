"""ctypes binding of libsavqa_b200.so -- the C ABI declared in include/savqa_b200.h.

There is deliberately NO fallback: if the shared library is missing or the device is not a B200 the ops raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# SAVQA_LIB: another build of the SAME library (kernel experiments, tools/build_variants.sh); never a different backend
LIB_PATH = os.environ.get("SAVQA_LIB") or os.path.join(os.path.dirname(_PKG), "lib", "libsavqa_b200.so")

i64 = C.c_int64
vp = C.c_void_p


class GemmEpilogue(C.Structure):
    """savqa_gemm_epilogue_t"""
    _fields_ = [("alpha", C.c_float), ("relu", C.c_int), ("accumulate", C.c_int), ("rowtab_period", C.c_int),
                ("bias", vp), ("res", vp), ("ld_res", i64), ("rowtab", vp), ("ld_rowtab", i64),
                ("gate_bf16", vp), ("ld_gate", i64), ("out_f32", vp), ("ld_out_f32", i64),
                ("out_bf16", vp), ("ld_out_bf16", i64), ("colsum", vp)]


class GemmProblem(C.Structure):
    """savqa_gemm_problem_t"""
    _fields_ = [("A", vp), ("lda", i64), ("B", vp), ("ldb", i64), ("M", C.c_int), ("K", C.c_int), ("epilogue", GemmEpilogue)]


class AttnArgs(C.Structure):
    """savqa_attn_args_t"""
    _fields_ = [("q", vp), ("ldq", i64), ("k", vp), ("ldk", i64), ("v", vp), ("ldv", i64),
                ("graph", vp), ("graph_n_stride", i64), ("graph_q_stride", i64),
                ("key_on", vp), ("query_on", vp),
                ("N", C.c_int), ("H", C.c_int), ("Tq", C.c_int), ("Tk", C.c_int), ("d", C.c_int),
                ("causal", C.c_int), ("renorm", C.c_int), ("engine", C.c_int),
                ("out", vp), ("ldo", i64), ("att", vp),
                ("dout", vp), ("ld_dout", i64), ("dq", vp), ("ld_dq", i64), ("dk", vp), ("ld_dk", i64),
                ("dv", vp), ("ld_dv", i64), ("scratch", vp), ("dbq", vp), ("dbk", vp), ("dbv", vp),
                ("graph_bits", vp), ("bits_n_stride", i64), ("bits_q_stride", i64), ("stats", vp), ("scale_d", C.c_int)]


class RowLnArgs(C.Structure):
    """savqa_rowln_args_t"""
    _fields_ = [("A", vp), ("lda", i64), ("B", vp), ("ldb", i64), ("b_mn_major", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("mode", C.c_int), ("relu", C.c_int),
                ("bias", vp), ("rowscale", vp), ("res", vp), ("ld_res", i64), ("gate_bf16", vp), ("ld_gate", i64),
                ("gamma", vp), ("beta", vp), ("eps", C.c_float),
                ("act_bf16", vp), ("ld_act", i64), ("pre", vp), ("ld_pre", i64), ("y", vp), ("ld_y", i64), ("y_bf16", vp), ("ld_yb", i64),
                ("on", vp), ("stats", vp), ("dxg_bf16", vp), ("ld_dxg", i64), ("dgamma", vp), ("dbeta", vp), ("dxsum", vp)]


#: every symbol include/savqa_b200.h declares: name -> argtypes (restype is int unless noted)
SIGNATURES = {
    "savqa_abi_version": [],
    "savqa_last_error": [],
    "savqa_device_check": [C.POINTER(C.c_int)],
    "savqa_launch_counts": [C.POINTER(i64), C.c_int],
    "savqa_build_masks": [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp],
    "savqa_build_masks_compact": [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp],
    "savqa_mil_nce_fwd": [vp, i64, vp, i64, vp, vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp],
    "savqa_mil_nce_bwd": [vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, i64, vp, i64, vp],
    "savqa_pack_graph_bits": [vp, i64, C.c_int, vp, C.c_int, vp],
    "savqa_gather_rows": [vp, i64, C.c_int, vp, i64, C.c_float, vp, i64, vp, i64, C.c_int, vp],
    "savqa_scatter_add_rows": [vp, i64, C.c_int, vp, i64, vp, i64, C.c_float, i64, vp],
    "savqa_scatter_add_rows_q48": [vp, i64, C.c_int, vp, i64, vp, i64, C.c_float, i64, vp],
    "savqa_cast_bf16": [vp, i64, vp, i64, i64, C.c_int, C.c_int, vp],
    "savqa_cast_transpose_bf16": [vp, i64, vp, i64, i64, C.c_int, C.c_int, vp],
    "savqa_row_nonzero": [vp, i64, i64, C.c_int, vp, vp, i64, vp],
    "savqa_relu_gate_bf16": [vp, C.c_int, i64, vp, i64, vp, i64, i64, C.c_int, i64, i64, vp],
    "savqa_fill_zero": [vp, i64, C.c_int, vp],
    "savqa_regroup_cols": [vp, i64, vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp],
    "savqa_colsum_bf16": [vp, i64, i64, C.c_int, vp, vp],
    "savqa_residual_layernorm_fwd": [vp, vp, vp, vp, C.c_float, i64, C.c_int, vp, vp, vp, vp, vp, vp],
    "savqa_layernorm_bwd": [vp, vp, vp, C.c_float, i64, C.c_int, vp, vp, vp, vp, vp, vp, vp],
    "savqa_gemm_bf16": [vp, i64, C.c_int, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(GemmEpilogue), C.c_int, vp],
    "savqa_gemm_bf16_grouped": [C.POINTER(GemmProblem), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp],
    "savqa_gemm_rowln": [C.POINTER(RowLnArgs), vp],
    "savqa_set_gemm_sm_limit": [C.c_int],
    "savqa_graph_attn_fwd": [C.POINTER(AttnArgs), vp],
    "savqa_graph_attn_bwd": [C.POINTER(AttnArgs), vp],
    "savqa_answer_loss": [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, vp, vp, vp],
    "savqa_adam_rows": [vp, vp, C.c_int, vp, vp, vp, i64, C.c_int, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, vp, C.c_int, vp],
    "savqa_adam_advance": [vp, C.c_float, C.c_float, C.c_float, vp],
    "savqa_adam_step": [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, vp, vp, vp],
}

_lib: Optional[C.CDLL] = None
_device_ok = False
launch_count = 0  # kernels-launching C-ABI calls made so far (bench.py reports it as gpu_launches)


def load() -> C.CDLL:
    """dlopen the library and bind every symbol; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"savqa_b200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "savqa_last_error" else C.c_int
    _lib = lib
    return lib


def require_device() -> None:
    """The product path runs on a B200 or not at all."""
    global _device_ok
    if _device_ok:
        return
    lib = load()
    if not torch.cuda.is_available():
        raise RuntimeError("savqa_b200: no CUDA device; the sm_100a kernels have no CPU fallback")
    sms = C.c_int(0)
    rc = lib.savqa_device_check(C.byref(sms))
    if rc != 0:
        raise RuntimeError("savqa_b200: " + lib.savqa_last_error().decode())
    _device_ok = True


LAUNCH_KINDS = ("gemm_pair", "gemm_single", "attn_fwd_tc", "attn_bwd_tc_shared", "attn_bwd_tc", "attn_fwd_simt", "attn_bwd_simt",
                "attn_row1_fwd", "attn_row1_bwd", "rowln_gemm", "mil_nce")


def launch_counts() -> dict:
    """Cumulative launches per kernel family (savqa_launch_counts): which engine did a shape take?"""
    buf = (i64 * len(LAUNCH_KINDS))()
    load().savqa_launch_counts(buf, len(LAUNCH_KINDS))
    return {k: int(buf[i]) for i, k in enumerate(LAUNCH_KINDS)}


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("savqa_b200: " + load().savqa_last_error().decode())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def call(name: str, *args) -> None:
    """Invoke an entry point on torch's current stream and raise on a non-zero return code."""
    global launch_count
    require_device()
    launch_count += 1
    check(getattr(_lib, name)(*args, stream_ptr()))
