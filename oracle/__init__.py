"""
ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithms for the hot path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import anything from this package; the product (`whisper-timestamped_b200/`)
never does, and fails loudly when its CUDA extension is missing.

Parity status: PARITY UNPINNED for the numerical core — the reference ships no golden
vectors for DTW / attention post-processing / mel / model forward and its
dependencies `openai-whisper` and `dtw-python` cannot be installed in this image, so the
oracle is pinned by (i) brute-force optimal-path enumeration and hand-computed tie cases for
DTW, (ii) the installed `scipy.ndimage.median_filter` as the importable sub-oracle for the
median filter, (iii) `transformers`' Whisper implementation as an independent cross-check of
the model forward and log-mel, (iv) the reference's own known-answer vectors for
`split_tokens_on_spaces` (tests/test_transcribe.py:722-902) replayed through a stub tokenizer.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) into oracle/_build/liboracle.so."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "dtw_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_dtw.restype = ctypes.c_int
        _LIB.oracle_dtw_batch_f32.restype = ctypes.c_int
    return _LIB


def dtw_symmetric1(lm: np.ndarray):
    """dtw.dtw(lm, step_pattern=symmetric1) → (index1s, index2s, jumps, distance).

    Follows transcribe.py:1581 + 1648-1652 (see dtw_oracle.c for the restated algorithm)."""
    lm = np.ascontiguousarray(lm, dtype=np.float64)
    n, m = lm.shape
    i1 = np.empty(n + m, dtype=np.int32)
    i2 = np.empty(n + m, dtype=np.int32)
    jumps = np.empty(n + 1, dtype=np.int32)
    dist = ctypes.c_double()
    ln = lib().oracle_dtw(
        lm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(m),
        i1.ctypes.data_as(ctypes.c_void_p), i2.ctypes.data_as(ctypes.c_void_p),
        jumps.ctypes.data_as(ctypes.c_void_p), ctypes.byref(dist))
    if ln < 0:
        raise ValueError("No warping path found")
    return i1[:ln].copy(), i2[:ln].copy(), jumps, dist.value


def dtw_fill(lm: np.ndarray):
    """Return (cm, sm) of the fill step — used by tests to check tie-breaking."""
    lm = np.ascontiguousarray(lm, dtype=np.float64)
    n, m = lm.shape
    cm = np.empty((n, m), dtype=np.float64)
    sm = np.empty((n, m), dtype=np.int32)
    lib().oracle_dtw_fill(lm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(m),
                          cm.ctypes.data_as(ctypes.c_void_p), sm.ctypes.data_as(ctypes.c_void_p))
    return cm, sm


def dtw_batch_f32(cost: np.ndarray, off, T, F):
    """Batched jumps for float32 cost matrices stored back to back (bench cpu_baseline)."""
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    T = np.ascontiguousarray(T, dtype=np.int32)
    F = np.ascontiguousarray(F, dtype=np.int32)
    joff = np.zeros(len(T), dtype=np.int64)
    joff[1:] = np.cumsum(T[:-1].astype(np.int64) + 1)
    jumps = np.empty(int((T.astype(np.int64) + 1).sum()), dtype=np.int32)
    rc = lib().oracle_dtw_batch_f32(
        cost.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
        T.ctypes.data_as(ctypes.c_void_p), F.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int(len(T)), jumps.ctypes.data_as(ctypes.c_void_p),
        joff.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("No warping path found")
    return jumps, joff
