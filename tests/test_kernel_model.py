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


@pytest.mark.parametrize("geo", [(16, 1), (8, 3), (32, 1)])
@pytest.mark.parametrize("late", [False, True])
def test_small_kernel_model_matches_oracle(late, geo):
    """Index arithmetic of dtw_small_kernel (3-tile ring + mirror, per-tile read base, bulk copies landing early or
    at the last moment) on negative costs incl. heavy ties, for shapes around every tile / pitch boundary."""
    from dtw_kernel_model import model_dtw_small
    rng = np.random.default_rng(5)
    shapes = [(1, 1), (1, 5), (2, 3), (3, 31), (24, 300), (31, 354), (31, 32), (17, 33), (8, 63), (9, 64), (10, 65), (30, 96),
              (31, 97), (12, 127), (13, 128), (14, 129), (24, 191), (5, 193), (31, 288), (2, 353)]
    shapes += [(int(rng.integers(1, 32)), int(rng.integers(1, 355))) for _ in range(25)]
    for n, (T, F) in enumerate(shapes):
        if n % 3 == 0:
            c = -(rng.random((T, F)).astype(np.float32) + np.float32(1e-3))
        elif n % 3 == 1:
            c = -np.ones((T, F), np.float32)
        else:
            c = -rng.integers(1, 4, (T, F)).astype(np.float32)
        _, _, jumps, _ = oracle.dtw_symmetric1(c.astype(np.float64))
        assert np.array_equal(jumps, model_dtw_small(c, late=late, TC=geo[0], LA=geo[1])), (T, F, late, geo)


def test_mma_k_permutation_model():
    """The permuted-k staging of lean_mma_kernel: with activations stored as word 4 s + q <- k 8 q + 2 s + {0, 1} per
    32-k block and every thread's weight fragment taken from ONE 16-byte load of 8 consecutive k, the two MMAs of a
    block contract exactly the 32 products of that block."""
    from dtw_kernel_model import mma_model
    rng = np.random.default_rng(3)
    for K in (32, 96, 384, 1280):
        x = rng.standard_normal((16, K))
        w = rng.standard_normal((8, K))
        assert np.allclose(mma_model(x, w), x @ w.T, rtol=1e-12, atol=1e-9), K


@pytest.mark.parametrize("NC,G", [(2, 1), (4, 1), (2, 2), (4, 2), (2, 4)])
@pytest.mark.parametrize("TR", [8, 16, 24, 32])
def test_lane_kernel_model_matches_oracle(TR, NC, G):
    """Recurrence order of dtw_lane_kernel (NC columns skewed by one row each in one lane's registers, G row bands per
    matrix running one column group apart and handing their last row down), its direction-word layout and backtrack, with garbage in the cells outside the matrix — on negative costs incl. heavy ties."""
    from dtw_kernel_model import model_dtw_lane
    rng = np.random.default_rng(11 + TR)
    shapes = [(1, 1), (1, 2), (1, 17), (2, 2), (2, 3), (TR, 16), (TR, 17), (TR, 31), (TR, 32), (TR, 33), (TR - 1, 48),
              (min(TR, 24), 300), (3, 8), (3, 9), (TR, 7), (TR, 8)]
    shapes += [(int(rng.integers(1, TR + 1)), int(rng.integers(1, 120))) for _ in range(12)]
    for n, (T, F) in enumerate(shapes):
        if n % 3 == 0:
            c = -(rng.random((T, F)).astype(np.float32) + np.float32(1e-3))
        elif n % 3 == 1:
            c = -np.ones((T, F), np.float32)
        else:
            c = -rng.integers(1, 4, (T, F)).astype(np.float32)
        _, _, jumps, _ = oracle.dtw_symmetric1(c.astype(np.float64))
        assert np.array_equal(jumps, model_dtw_lane(c, TR=TR, NC=NC, G=G)), (T, F, TR, NC, G)
        if n % 4 == 0:
            assert np.array_equal(jumps, model_dtw_lane(c, TR=TR, NC=NC, G=G, stale=np.float32("nan"))), (T, F, TR, NC, G, "nan")


@pytest.mark.parametrize("NC,G", [(2, 1), (4, 1), (2, 2), (4, 2), (2, 4)])
def test_lane_kernel_staging_ring_schedule(NC, G):
    """Slot reuse of the lane kernel's two-tile cp.async ring: whether a copy lands at once or only at the next wait, every
    band reads its own tile (a step's shared-memory reads precede the copies it issues, in program order)."""
    from dtw_kernel_model import lane_ring_schedule_ok
    for ntile in (1, 2, 3, 7, 38):
        assert lane_ring_schedule_ok(NC, G, ntile, early=True), (NC, G, ntile, "early")
        assert lane_ring_schedule_ok(NC, G, ntile, early=False), (NC, G, ntile, "late")
    if G > 2:
        # two steps earlier the last band is still two groups inside the tile being replaced
        assert not lane_ring_schedule_ok(NC, G, 7, early=True, issue_step=G - 3), (NC, G, "two steps too early must fail")
