"""ctypes binding of libwts.so — the C-ABI declared in include/wts.h.

PyTorch tensors are only the memory carrier: every call passes raw device pointers, sizes and
the current CUDA stream.  There is NO CPU fallback: if the library is missing or cannot be
loaded the import of the product fails loudly.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwts.so")


class WtsError(RuntimeError):
    pass


class SegDesc(ctypes.Structure):
    """Mirror of `WtsSegDesc` (include/wts.h)."""
    _fields_ = [
        ("window", ctypes.c_int32), ("row0", ctypes.c_int32), ("last_row", ctypes.c_int32),
        ("T", ctypes.c_int32), ("f0", ctypes.c_int32), ("F", ctypes.c_int32),
        ("max_dur", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("cost_off", ctypes.c_int64), ("jumps_off", ctypes.c_int64),
        ("dir_off", ctypes.c_int64), ("bnd_off", ctypes.c_int64),
    ]


SEG_DTYPE = np.dtype([
    ("window", "<i4"), ("row0", "<i4"), ("last_row", "<i4"), ("T", "<i4"), ("f0", "<i4"),
    ("F", "<i4"), ("max_dur", "<i4"), ("flags", "<i4"), ("cost_off", "<i8"), ("jumps_off", "<i8"),
    ("dir_off", "<i8"), ("bnd_off", "<i8")])
assert SEG_DTYPE.itemsize == ctypes.sizeof(SegDesc) == 64


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python whisper-timestamped_b200/build.py` "
            "(the product has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.wts_version.restype = ctypes.c_int
    lib.wts_last_error.restype = ctypes.c_char_p
    lib.wts_dtw_dir_words.restype = ctypes.c_int64
    lib.wts_dtw_dir_words.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.wts_dtw_bnd_doubles.restype = ctypes.c_int64
    lib.wts_dtw_bnd_doubles.argtypes = [ctypes.c_int32, ctypes.c_int32]
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    lib.wts_attn_prep_batch.restype = ctypes.c_int
    lib.wts_attn_prep_batch.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]
    lib.wts_dtw_batch.restype = ctypes.c_int
    lib.wts_dtw_batch.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    return lib


lib = _load()

EXPORTED_SYMBOLS = [
    "wts_version", "wts_last_error", "wts_dtw_dir_words", "wts_dtw_bnd_doubles",
    "wts_attn_prep_batch", "wts_dtw_batch",
]


def check(rc: int, what: str):
    if rc != 0:
        raise WtsError(f"{what} failed ({rc}): {lib.wts_last_error().decode(errors='replace')}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    if t is None:
        return None
    assert t.is_contiguous(), "libwts needs contiguous buffers"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise WtsError(f"{name} must live on a CUDA device: libwts has no CPU path")
