"""ctypes binding of libarcface_b200.so (the C ABI declared in include/arcface_b200.h).

There is no fallback: if the shared library is missing, or the device is not a B200-class GPU
(compute capability 10.x), every entry point raises.  The library is built in-tree by
`make -C multimodalsimilar_b200/csrc` (see __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libarcface_b200.so")
# Diagnostic build (make DIAG=1): the same kernels plus the ARCFACE_B200_* environment knobs and the in-kernel wait
# profiler.  Only tools/ ask for it; tests, bench.py and the product load the release library.
if os.environ.get("ARCFACE_B200_DIAG") == "1":
    LIB_PATH = os.environ.get("ARCFACE_B200_DIAG_LIB") or os.path.join(_HERE, "libarcface_b200_diag.so")

OK = 0
ERROR_NAMES = {-1: "E_ARCH", -2: "E_SHAPE", -3: "E_LAYOUT", -4: "E_WORKSPACE", -5: "E_CUDA", -6: "E_ARG"}
MAX_BATCH = 2048

# name -> (restype, argtypes).  Kept in the order of include/arcface_b200.h; tests/test_abi.py checks
# that every symbol the header declares is exported and listed here.
SIGNATURES = {
    "arcface_b200_version": (c_int32, [POINTER(c_int32), POINTER(c_int32)]),
    "arcface_b200_last_error": (c_char_p, []),
    "arcface_b200_device_ok": (c_int32, []),
    "arcface_b200_normalize_cast": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "arcface_b200_normalize_cast3": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "arcface_b200_normalize_cast_gather": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_accumulate": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "arcface_b200_scatter_rows": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p]),
    "arcface_b200_label_margin": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_int64,
         c_float, c_float, c_float, c_float, c_float, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "arcface_b200_forward_parts": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_int32)]),
    "arcface_b200_forward_stats": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_float, c_void_p, c_void_p, c_void_p,
         c_int32, c_void_p],
    ),
    "arcface_b200_forward_fused_workspace_bytes": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_size_t)]),
    "arcface_b200_forward_stats_fused": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_int32, c_void_p, c_size_t, c_void_p],
    ),
    "arcface_b200_combine_partials": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_finalize_rows": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                  c_void_p, c_void_p, c_void_p]),
    "arcface_b200_finalize_rows_strided": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p,
                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_logits": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_float, c_void_p, c_int64, c_void_p]),
    "arcface_b200_topk_workspace_bytes": (c_int32, [c_int32, c_int32, c_int64, c_int32, POINTER(c_size_t)]),
    "arcface_b200_cosine_topk": (
        c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32, c_float, c_int64, c_void_p, c_void_p,
                  c_void_p, c_size_t, c_void_p]),
    "arcface_b200_topk_merge": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_backward_workspace_bytes": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_size_t)]),
    "arcface_b200_backward_plan": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_int64), POINTER(c_int32)]),
    "arcface_b200_backward_launches": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_int32)]),
    "arcface_b200_backward": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
         c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "arcface_b200_backward_prec": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
         c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int32, c_void_p],
    ),
    "arcface_b200_two_stream_concat": (
        c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_two_stream_concat_bwd": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_normalize_bwd_x": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "arcface_b200_scale_grads": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "arcface_b200_scale_copy": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "arcface_b200_pack_xy": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "arcface_b200_adamw_normalize": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_double, c_double, c_double, c_double,
                  c_double, c_int64, c_void_p, c_void_p, c_void_p]),
    "arcface_b200_p2p_exchange": (
        c_int32, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_int32, c_int32, c_size_t, c_int32, c_void_p,
                  c_void_p, c_void_p]),
    "arcface_b200_p2p_gather_split": (
        c_int32, [c_void_p, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_int32, c_int32, c_size_t, c_int32,
                  c_void_p, c_void_p, c_void_p]),
    "arcface_b200_normalize_bwd_x_sum": (
        c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "arcface_b200_step_workspace_bytes": (c_int32, [c_int32, c_int32, c_int64, POINTER(c_size_t)]),
    "arcface_b200_step_host": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_float, c_float, c_int32, c_float, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
}

_lib = None


class ArcfaceB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libarcface_b200: %s (%d): %s" % (ERROR_NAMES.get(code, "E_?"), code, message))
        self.code = code


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "multimodalsimilar_b200: %s is missing. Build it with `make -C multimodalsimilar_b200/csrc` "
                "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU / PyTorch fallback." % LIB_PATH
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().arcface_b200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code: int) -> None:
    if code != OK:
        raise ArcfaceB200Error(code, last_error())


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
