// Median of 9 by a fixed 19-exchange selection network, and the index map of
// scipy.ndimage's default boundary mode ('reflect' = half-sample symmetric, d c b a | a b c d).
// Shared by the CUDA prep kernel and by a host-side unit test (tests/test_host_logic.py builds it
// with gcc), so the network is verified on the CPU against numpy before it ever runs on a GPU.
// Replaces scipy.ndimage.median_filter(w, (1, 1, 9)) at
// /root/reference/whisper_timestamped/transcribe.py:1546.
#pragma once
#if defined(__CUDACC__)
#define WTS_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define WTS_HD static inline
#endif

#define WTS_CE(a, b) { const float _lo = fminf(a, b); b = fmaxf(a, b); a = _lo; }

WTS_HD float wts_median9(float p0, float p1, float p2, float p3, float p4, float p5, float p6,
                         float p7, float p8)
{
    WTS_CE(p1, p2); WTS_CE(p4, p5); WTS_CE(p7, p8);
    WTS_CE(p0, p1); WTS_CE(p3, p4); WTS_CE(p6, p7);
    WTS_CE(p1, p2); WTS_CE(p4, p5); WTS_CE(p7, p8);
    WTS_CE(p0, p3); WTS_CE(p5, p8); WTS_CE(p4, p7);
    WTS_CE(p3, p6); WTS_CE(p1, p4); WTS_CE(p2, p5);
    WTS_CE(p4, p7); WTS_CE(p4, p2); WTS_CE(p6, p4);
    WTS_CE(p4, p2);
    return p4;
}

// Two adjacent medians of 9 share 8 of their samples: with s3 <= s4 the two middle order statistics of the
// shared samples x[c-3 .. c+4], median9(x[c-4 .. c+4]) = clamp(x[c-4], s3, s4) and
// median9(x[c-3 .. c+5]) = clamp(x[c+5], s3, s4).  17 compare-exchanges (Batcher's 8-sorter pruned to ranks
// 3 and 4, checked exhaustively with the 0-1 principle) + 2 clamps for TWO outputs.
WTS_HD void wts_median9_pair(const float* v /* 10 samples x[c-4 .. c+5] */, float* m0, float* m1)
{
    float a0 = v[1], a1 = v[2], a2 = v[3], a3 = v[4], a4 = v[5], a5 = v[6], a6 = v[7], a7 = v[8];
    WTS_CE(a0, a1); WTS_CE(a2, a3); WTS_CE(a0, a2); WTS_CE(a1, a3); WTS_CE(a1, a2);
    WTS_CE(a4, a5); WTS_CE(a6, a7); WTS_CE(a4, a6); WTS_CE(a5, a7); WTS_CE(a5, a6);
    WTS_CE(a0, a4); WTS_CE(a2, a6); WTS_CE(a2, a4); WTS_CE(a1, a5); WTS_CE(a3, a7); WTS_CE(a3, a5);
    WTS_CE(a3, a4);
    *m0 = fmaxf(a3, fminf(v[0], a4));
    *m1 = fmaxf(a3, fminf(v[9], a4));
}

// index of the sample that position p (may be < 0 or >= n) maps to under symmetric reflection
WTS_HD int wts_reflect_index(int p, int n)
{
    const int period = 2 * n;
    int q = p % period;
    if (q < 0) q += period;
    return q < n ? q : period - 1 - q;
}
