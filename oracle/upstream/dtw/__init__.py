"""ORACLE stand-in for `dtw-python` (see oracle/upstream/README.md).  Only what the reference
calls at transcribe.py:1572-1581 / 1598 / 1648-1652: dtw.dtw(x, step_pattern=...) -> object with
.index1s / .index2s (and .index1 / .index2 / .distance)."""
import numpy as np

from . import stepPattern  # noqa: F401


class DTW:
    pass


def dtw(x, y=None, dist_method="euclidean", step_pattern="symmetric2", window_type=None, window_args={},
        keep_internals=False, distance_only=False, open_end=False, open_begin=False):
    import oracle
    assert y is None, "oracle dtw: only the precomputed local-cost form is restated"
    sp = step_pattern
    if isinstance(sp, str):
        sp = getattr(stepPattern, sp)
    if sp is not stepPattern.symmetric1:
        raise NotImplementedError("oracle dtw: only symmetric1 is restated")
    i1, i2, _, dist = oracle.dtw_symmetric1(np.asarray(x, dtype=np.float64))
    out = DTW()
    out.index1 = out.index1s = i1.astype(int)
    out.index2 = out.index2s = i2.astype(int)
    out.distance = dist
    out.stepPattern = sp
    out.N, out.M = np.asarray(x).shape
    return out
