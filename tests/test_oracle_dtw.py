"""Pins the DTW oracle (oracle/dtw_oracle.c).  PARITY UNPINNED against dtw-python itself (not
installable here, no reference fixture — SURVEY.md §8c); pinned instead by brute force over all
monotone paths, by hand-computed tie cases, and by the committed golden vectors."""
import itertools
import os
import sys

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def all_paths(n, m):
    """Every symmetric1 path (0,0)->(n-1,m-1) as a list of cells."""
    def rec(i, j):
        if i == 0 and j == 0:
            yield [(0, 0)]
            return
        for di, dj in ((1, 1), (0, 1), (1, 0)):
            ii, jj = i - di, j - dj
            if ii >= 0 and jj >= 0:
                for p in rec(ii, jj):
                    yield p + [(i, j)]
    return rec(n - 1, m - 1)


def path_cost(lm, path):
    c = lm[path[0]]
    for cell in path[1:]:
        c = c + lm[cell]        # same association as cm[pred] + lm[i,j]
    return c


@pytest.mark.parametrize("n,m", [(1, 1), (1, 4), (2, 2), (2, 5), (3, 3), (3, 5), (4, 5), (4, 6)])
def test_bruteforce_optimum(n, m):
    rng = np.random.default_rng(n * 100 + m)
    for trial in range(20):
        lm = rng.standard_normal((n, m)) if trial % 2 else -rng.random((n, m)).astype(np.float32).astype(np.float64)
        i1, i2, jumps, dist = oracle.dtw_symmetric1(lm)
        best = min(path_cost(lm, p) for p in all_paths(n, m))
        assert dist == best                              # exact: rounding is monotone
        path = list(zip(i1.tolist(), i2.tolist()))
        assert path[0] == (0, 0) and path[-1] == (n - 1, m - 1)
        assert path_cost(lm, path) == best
        for (a, b), (c, d) in zip(path[:-1], path[1:]):
            assert (c - a, d - b) in ((1, 1), (0, 1), (1, 0))


def test_tie_break_constant_matrix():
    # all-equal costs: every tie must go to the diagonal (pattern 1 before 2 before 3)
    n, m = 4, 9
    lm = np.ones((n, m))
    cm, sm = oracle.dtw_fill(lm)
    for i in range(n):
        for j in range(m):
            assert cm[i, j] == max(i, j) + 1
            exp = 0 if (i, j) == (0, 0) else 2 if i == 0 else 3 if j == 0 else 1
            # interior: diag ties with left (j>i) or with up (j<i) or is strictly best (j==i) -> diag
            assert sm[i, j] == exp, (i, j, sm[i, j])
    i1, i2, jumps, _ = oracle.dtw_symmetric1(lm)
    assert jumps.tolist() == [0, m - n + 1, m - n + 2, m - n + 3, m - 1]


def test_tie_break_left_before_up():
    # cell (1,1): make diag expensive and left == up -> must choose left (pattern 2)
    lm = np.array([[0.0, 1.0], [1.0, 0.0]])
    lm[0, 0] = 5.0      # cm: [[5,6],[6,?]] ; candidates diag=5, left=6, up=6 -> diag
    cm, sm = oracle.dtw_fill(lm)
    assert sm[1, 1] == 1
    lm2 = np.array([[0.0, -3.0], [-3.0, 0.0]])   # cm00=0, cm01=-3, cm10=-3: left == up < diag
    cm, sm = oracle.dtw_fill(lm2)
    assert sm[1, 1] == 2


def test_rounding_tie_needs_add_before_compare():
    # predecessors differ by less than half an ulp of the sum: sums tie, earlier pattern wins
    big = -1.0
    eps = 2.0 ** -60
    lm = np.array([[0.0, eps], [0.0, big]])
    # cm00=0, cm01=eps (left), cm10=0 (up).  cell(1,1): diag=0+big, left: cm10.. careful: p2=(1,0)
    cm, sm = oracle.dtw_fill(lm)
    # candidates: p1 cm[0,0]+big = -1 ; p2 cm[1,0]+big = -1 ; p3 cm[0,1]+big = fl(eps-1) = -1
    assert sm[1, 1] == 1 and cm[1, 1] == -1.0
    lm[0, 0] = eps * 4          # now p1 = fl(4eps - 1) = -1 as well -> still diag although cm00 > cm10
    cm, sm = oracle.dtw_fill(lm)
    assert sm[1, 1] == 1


def test_jumps_definition():
    rng = np.random.default_rng(7)
    lm = -rng.random((9, 40))
    i1, i2, jumps, _ = oracle.dtw_symmetric1(lm)
    for t in range(9):
        assert jumps[t] == i2[i1 == t].min()
    assert jumps[9] == 39


def test_golden_vectors():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    g = np.load(os.path.join(HERE, "golden", "dtw_golden.npz"))
    for n, (kind, c) in enumerate(make_golden.dtw_cases()):
        if f"dtw{n}_cost" in g:
            assert np.array_equal(g[f"dtw{n}_cost"], c)
        i1, i2, jumps, dist = oracle.dtw_symmetric1(c.astype(np.float64))
        assert np.array_equal(jumps, g[f"dtw{n}_jumps"])
        assert np.array_equal(i1, g[f"dtw{n}_i1"]) and np.array_equal(i2, g[f"dtw{n}_i2"])
        assert dist == float(g[f"dtw{n}_dist"])
