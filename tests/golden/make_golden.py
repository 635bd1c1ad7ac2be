"""Generates tests/golden/*.npz from the oracle (run once here; the vectors are committed).

The reference ships no golden vectors for the numerical core and its
dependencies are not installable in this image, so these are ORACLE outputs: seeded inputs,
oracle results.  They pin the oracle against accidental drift and give the GPU tests fixtures
that do not need the oracle's C library.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle.prep import attn_cost  # noqa: E402


def dtw_cases():
    rng = np.random.default_rng(1234)
    cases = []
    shapes = [(1, 1), (1, 7), (2, 2), (3, 5), (4, 5), (7, 7), (12, 150), (24, 300), (31, 64),
              (32, 64), (33, 90), (62, 63), (63, 200), (100, 101), (224, 1500)]
    for (T, F) in shapes:
        for kind in ("uniform", "ties_const", "ties_int", "checker", "normal"):
            if kind == "uniform":
                c = -rng.random((T, F), dtype=np.float32)
            elif kind == "ties_const":
                c = -np.ones((T, F), np.float32)
            elif kind == "ties_int":
                c = -rng.integers(0, 3, (T, F)).astype(np.float32)
            elif kind == "checker":
                c = -((np.add.outer(np.arange(T), np.arange(F)) % 2).astype(np.float32))
            else:
                c = rng.standard_normal((T, F)).astype(np.float32)
            cases.append((kind, c))
    return cases


def main():
    out = {}
    for n, (kind, c) in enumerate(dtw_cases()):
        i1, i2, jumps, dist = oracle.dtw_symmetric1(c.astype(np.float64))
        T, F = c.shape
        if T * F <= 64 * 256:
            out[f"dtw{n}_cost"] = c
        else:   # big ones are regenerated from the seed by the test; keep only the answers
            out[f"dtw{n}_shape"] = np.array([T, F])
        out[f"dtw{n}_kind"] = np.array(kind)
        out[f"dtw{n}_jumps"] = jumps
        out[f"dtw{n}_i1"] = i1.astype(np.int16)
        out[f"dtw{n}_i2"] = i2.astype(np.int16)
        out[f"dtw{n}_dist"] = np.array(dist)
    np.savez_compressed(os.path.join(HERE, "dtw_golden.npz"), **out)

    rng = np.random.default_rng(4321)
    pout = {}
    specs = [(10, 12, 150, 40, 0), (8, 24, 300, 0, 0), (6, 5, 7, 3, 0), (10, 3, 3, 0, 0),
             (8, 20, 120, 1300, 1350), (10, 40, 90, 100, 150), (10, 2, 9, 0, 0),
             (8, 20, 120, 10, 60), (6, 9, 200, 0, 1), (10, 12, 80, 70, 70)]
    for n, (N, T, F, f0, max_dur) in enumerate(specs):
        qk = (3.0 * rng.standard_normal((N, T, 1500))).astype(np.float32)
        ridge = 6.0 * np.exp(-((np.arange(1500)[None, :] - (f0 + F * (np.arange(T)[:, None] + 0.5) / T)) / 8.0) ** 2)
        qk += ridge.astype(np.float32)[None]
        cost = attn_cost(qk, f0, f0 + F, max_duration=max_dur or None)
        _, _, jumps, _ = oracle.dtw_symmetric1(cost)
        pout[f"prep{n}_qk"] = qk[:, :, f0:f0 + F].copy()
        pout[f"prep{n}_spec"] = np.array([N, T, F, f0, max_dur])
        pout[f"prep{n}_cost"] = cost.astype(np.float32)
        assert np.array_equal(cost.astype(np.float32).astype(np.float64), cost)
        pout[f"prep{n}_jumps"] = jumps
    np.savez_compressed(os.path.join(HERE, "prep_golden.npz"), **pout)
    print("written", len(out), len(pout))


if __name__ == "__main__":
    main()
