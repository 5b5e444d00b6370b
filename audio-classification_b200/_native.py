"""ctypes binding of ``libb200sep.so`` (C ABI in ``include/b200sep.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``csrc/build.py``).  There is no
Python/CPU fallback: if the shared object is missing or does not load, importing the product API
raises ``RuntimeError`` (the reference's own wrapper makes the same promise,
``/root/reference/src/osd/separation.py:1-6``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SEP_LIB") or os.path.join(_HERE, "libb200sep.so")  # env override: kernel-variant A/B runs

PREC_FP32, PREC_FP16_TC, PREC_FP16_SIMT = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "fp16": PREC_FP16_TC, "fp16_tc": PREC_FP16_TC,
              "fp16_simt": PREC_FP16_SIMT}
MASK_ACTS = {"sigmoid": 0, "relu": 1}
PIT_VARIANTS = {"overlap3": 0, "evalsrc": 1}

# every symbol include/b200sep.h declares (tests check the .so exports all of them)
EXPORTS = [
    "b200sep_version", "b200sep_lpt_partition", "b200sep_last_error", "b200sep_model_create", "b200sep_model_destroy",
    "b200sep_workspace_bytes", "b200sep_separate_batch", "b200sep_pit_workspace_bytes", "b200sep_pit_workspace_bytes_n",
    "b200sep_pit_sisdr_batch", "b200sep_debug_gemm", "b200sep_debug_set_stop",
    "b200sep_debug_layout", "b200sep_profile_enable", "b200sep_profile_read",
    "b200sep_pipeline", "b200sep_debug_timeline", "b200sep_pit_profile_enable", "b200sep_pit_profile_read",
    "b200sep_host_workspace_bytes", "b200sep_separate_batch_host", "b200sep_split_count", "b200sep_set_max_split",
    "b200sep_resample_out_lens", "b200sep_resample_batch", "b200sep_frame_rms_out_lens",
    "b200sep_frame_rms_batch", "b200sep_gt_overlap_mask_batch", "b200sep_debug_ts_probe",
]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_src", "n_filters", "kernel_size", "stride", "bn_chan", "hid_chan", "skip_chan",
        "conv_kernel_size", "n_blocks", "n_repeats", "mask_act", "precision")]


_FP = C.POINTER(C.c_float)
_FPP = C.POINTER(_FP)


class Weights(C.Structure):
    _fields_ = [
        ("enc_filters", _FP), ("dec_filters", _FP), ("bn_gamma", _FP), ("bn_beta", _FP),
        ("bn_weight", _FP), ("bn_bias", _FP),
        ("in_weight", _FPP), ("in_bias", _FPP), ("prelu1", _FPP), ("norm1_gamma", _FPP),
        ("norm1_beta", _FPP), ("dw_weight", _FPP), ("dw_bias", _FPP), ("prelu2", _FPP),
        ("norm2_gamma", _FPP), ("norm2_beta", _FPP), ("res_weight", _FPP), ("res_bias", _FPP),
        ("skip_weight", _FPP), ("skip_bias", _FPP),
        ("mask_prelu", _FP), ("mask_weight", _FP), ("mask_bias", _FP), ("gln_eps", _FP),
    ]


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """The loaded native library; raises RuntimeError (never falls back) when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"b200sep native library not found at {LIB_PATH}; build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback.")
        try:
            L = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"failed to load {LIB_PATH}: {e}") from e
        vp, i32, i64p, i32p = C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        L.b200sep_version.restype = C.c_int
        L.b200sep_lpt_partition.restype = C.c_int
        L.b200sep_lpt_partition.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.b200sep_last_error.restype = C.c_char_p
        L.b200sep_model_create.restype = C.c_int
        L.b200sep_model_create.argtypes = [C.POINTER(Config), C.POINTER(Weights), C.c_int,
                                           C.POINTER(vp)]
        L.b200sep_model_destroy.restype = None
        L.b200sep_model_destroy.argtypes = [vp]
        L.b200sep_workspace_bytes.restype = C.c_size_t
        L.b200sep_workspace_bytes.argtypes = [vp, i32p, i32]
        L.b200sep_separate_batch.restype = C.c_int
        L.b200sep_separate_batch.argtypes = [vp, vp, i64p, i32p, i32, vp, vp, C.c_size_t, vp, vp]
        L.b200sep_resample_out_lens.restype = C.c_int
        L.b200sep_resample_out_lens.argtypes = [i32p, i32, i32, i32, i32, i32p]
        L.b200sep_resample_batch.restype = C.c_int
        L.b200sep_resample_batch.argtypes = [vp, i32p, i32, i32, i32, i32, vp, vp]
        L.b200sep_frame_rms_out_lens.restype = C.c_int
        L.b200sep_frame_rms_out_lens.argtypes = [i32p, i32, i32, i32, i32p]
        L.b200sep_frame_rms_batch.restype = C.c_int
        L.b200sep_frame_rms_batch.argtypes = [vp, i32p, i32, i32, i32, vp, vp]
        L.b200sep_gt_overlap_mask_batch.restype = C.c_int
        L.b200sep_gt_overlap_mask_batch.argtypes = [vp, vp, i32p, i32, C.c_float, vp, vp]
        L.b200sep_split_count.restype = C.c_int
        L.b200sep_split_count.argtypes = [vp, i32p, i32]
        L.b200sep_set_max_split.restype = C.c_int
        L.b200sep_set_max_split.argtypes = [vp, i32]
        L.b200sep_host_workspace_bytes.restype = C.c_size_t
        L.b200sep_host_workspace_bytes.argtypes = [vp, i32p, i32]
        L.b200sep_separate_batch_host.restype = C.c_int
        L.b200sep_separate_batch_host.argtypes = [vp, C.POINTER(vp), i32p, i32, vp, vp, C.c_size_t, vp, vp]
        L.b200sep_pit_workspace_bytes.restype = C.c_size_t
        L.b200sep_pit_workspace_bytes.argtypes = [i32p, i32]
        L.b200sep_pit_workspace_bytes_n.restype = C.c_size_t
        L.b200sep_pit_workspace_bytes_n.argtypes = [i32p, i32, i32]
        L.b200sep_pit_sisdr_batch.restype = C.c_int
        L.b200sep_pit_sisdr_batch.argtypes = [vp, vp, vp, i64p, i32p, i32, i32, i32, i32, vp, vp,
                                              vp, vp, vp, vp, C.c_size_t, vp]
        L.b200sep_debug_gemm.restype = C.c_int
        L.b200sep_debug_gemm.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
        L.b200sep_debug_ts_probe.restype = C.c_int
        L.b200sep_debug_ts_probe.argtypes = [vp, vp, vp, i32, vp]
        L.b200sep_debug_set_stop.restype = C.c_int
        L.b200sep_debug_set_stop.argtypes = [vp, C.c_int]
        L.b200sep_debug_layout.restype = C.c_int
        L.b200sep_debug_layout.argtypes = [vp, i32p, i32, C.POINTER(C.c_uint64)]
        L.b200sep_profile_enable.restype = C.c_int
        L.b200sep_profile_enable.argtypes = [vp, C.c_int]
        L.b200sep_profile_read.restype = C.c_int
        L.b200sep_profile_read.argtypes = [vp, C.POINTER(C.c_float), i32, i32p]
        L.b200sep_debug_timeline.restype = C.c_int
        L.b200sep_debug_timeline.argtypes = [vp]
        L.b200sep_pipeline.restype = C.c_int
        L.b200sep_pipeline.argtypes = [vp]
        _lib = L
    return _lib


def last_error() -> str:
    return lib().b200sep_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what}: {last_error()}")


MAX_SPLIT = 4          # B200SEP_MAX_SPLIT: h_flags words per call
FLAG_WAIT_TIMEOUT = 1  # B200SEP_FLAG_WAIT_TIMEOUT
FLAG_NONFINITE = 2     # B200SEP_FLAG_NONFINITE
FLAG_EPS_DOMINATED = 4  # B200SEP_FLAG_EPS_DOMINATED


def i32_array(vals):
    """ctypes int32 array of `vals` (through NumPy: a sweep passes tens of thousands of lengths per call)."""
    import numpy as np
    a = np.array(vals, dtype=np.int32).reshape(-1)  # (always a fresh, writable copy: from_buffer shares it)
    return (C.c_int32 * a.size).from_buffer(a)


def i64_array(vals):
    import numpy as np
    a = np.array(vals, dtype=np.int64).reshape(-1)
    return (C.c_int64 * a.size).from_buffer(a)
