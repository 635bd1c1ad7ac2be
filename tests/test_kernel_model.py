"""The lane-level numpy model of the CUDA DTW kernel (tests/dtw_kernel_model.py) must agree with
the oracle bit for bit — this is how the kernel's index arithmetic (strips, skewed staging,
direction packing, clz backtrack) is checked on the CPU-only build box."""
import numpy as np
import pytest

import oracle
from dtw_kernel_model import model_dtw


@pytest.mark.parametrize("seed", range(6))
def test_model_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    for trial in range(12):
        T = int(rng.integers(1, 70))
        F = int(rng.integers(T, T + 100))
        kind = trial % 4
        if kind == 0:
            c = -rng.random((T, F)).astype(np.float32)
        elif kind == 1:
            c = -np.ones((T, F), np.float32)
        elif kind == 2:
            c = -rng.integers(0, 3, (T, F)).astype(np.float32)
        else:
            c = rng.standard_normal((T, F)).astype(np.float32)
        _, _, jumps, _ = oracle.dtw_symmetric1(c.astype(np.float64))
        assert np.array_equal(jumps, model_dtw(c)), (seed, trial, T, F)


def test_model_exact_strip_boundaries():
    rng = np.random.default_rng(99)
    for T in (30, 31, 32, 61, 62, 63, 93):
        c = -rng.random((T, T + 17)).astype(np.float32)
        _, _, jumps, _ = oracle.dtw_symmetric1(c.astype(np.float64))
        assert np.array_equal(jumps, model_dtw(c)), T
