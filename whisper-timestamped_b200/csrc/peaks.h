// Peak analysis behind detect_disfluencies, host + device.
//
// The reference (/root/reference/whisper_timestamped/transcribe.py:1656-1683) looks, for every token, at the slice of
// its (negated) alignment-cost row between the token's two DTW jumps and calls
//     scipy.signal.find_peaks(x, width=3, prominence=0.02)
// When MORE THAN ONE peak survives, the token is taken to start at round(left_ips[-1]) — the left half-prominence
// crossing of the LAST peak — and whatever lies before is reported as a disfluency "[*]".
//
// This restates the four scipy stages that call goes through (local maxima with plateau midpoints, prominences over
// the whole slice, the prominence filter, widths at rel_height 0.5 with linear interpolation, the width filter) in
// float64, operation for operation, so that the rounded index is identical.  tests/test_host_logic.py compiles this
// header with gcc and checks it against scipy itself.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WTS_PK_HD __host__ __device__ __forceinline__
#else
#define WTS_PK_HD static inline
#endif

// x(i) = -(double)cost_row[begin + i],  0 <= i < n
WTS_PK_HD double wts_pk_x(const float* cost_row, int begin, int i) { return -(double)cost_row[begin + i]; }

// Returns round-half-even(left_ip of the last surviving peak) when at least two peaks survive, else -1.
WTS_PK_HD int wts_disfluency_left(const float* cost_row, int begin, int n, double min_prominence, double min_width)
{
    int survivors = 0;
    double last_left_ip = 0.0;
    int i = 1;
    const int i_last = n - 1;
    while (i < i_last) {
        const double xi = wts_pk_x(cost_row, begin, i);
        if (wts_pk_x(cost_row, begin, i - 1) < xi) {
            int ahead = i + 1;
            while (ahead < i_last && wts_pk_x(cost_row, begin, ahead) == xi) ++ahead;
            if (wts_pk_x(cost_row, begin, ahead) < xi) {
                const int peak = (i + (ahead - 1)) / 2;            // plateau midpoint
                const double xp = wts_pk_x(cost_row, begin, peak);
                // ---- prominence over the whole slice
                double left_min = xp, right_min = xp;
                int left_base = peak, right_base = peak;
                for (int k = peak; k >= 0 && wts_pk_x(cost_row, begin, k) <= xp; --k) {
                    const double v = wts_pk_x(cost_row, begin, k);
                    if (v < left_min) { left_min = v; left_base = k; }
                }
                for (int k = peak; k <= n - 1 && wts_pk_x(cost_row, begin, k) <= xp; ++k) {
                    const double v = wts_pk_x(cost_row, begin, k);
                    if (v < right_min) { right_min = v; right_base = k; }
                }
                const double prominence = xp - (left_min > right_min ? left_min : right_min);
                if (prominence >= min_prominence) {
                    // ---- width at half prominence, linear interpolation on both flanks
                    const double height = xp - prominence * 0.5;
                    int k = peak;
                    while (left_base < k && height < wts_pk_x(cost_row, begin, k)) --k;
                    double left_ip = (double)k;
                    {
                        const double xk = wts_pk_x(cost_row, begin, k);
                        if (xk < height) left_ip += (height - xk) / (wts_pk_x(cost_row, begin, k + 1) - xk);
                    }
                    k = peak;
                    while (k < right_base && height < wts_pk_x(cost_row, begin, k)) ++k;
                    double right_ip = (double)k;
                    {
                        const double xk = wts_pk_x(cost_row, begin, k);
                        if (xk < height) right_ip -= (height - xk) / (wts_pk_x(cost_row, begin, k - 1) - xk);
                    }
                    if (right_ip - left_ip >= min_width) {
                        ++survivors;
                        last_left_ip = left_ip;
                    }
                }
                i = ahead;
            }
        }
        ++i;
    }
    if (survivors < 2) return -1;
    return (int)rint(last_left_ip);
}
