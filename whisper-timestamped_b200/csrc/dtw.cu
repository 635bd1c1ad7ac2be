// Batched monotonic DTW (symmetric1) as a warp-wavefront kernel for sm_100a.
//
// Replaces  dtw.dtw(weights, step_pattern=dtw.stepPattern.symmetric1)   (T.py:1572-1581)
// and the jumps extraction                                              (T.py:1648-1652)
// where T.py = /root/reference/whisper_timestamped/transcribe.py.
//
// Semantics restated from dtw-python (see oracle/dtw_oracle.c): for every cell the three
// candidates  cm[i-1,j-1]+lm, cm[i,j-1]+lm, cm[i-1,j]+lm  are formed in float64 (one add each) and
// the FIRST strict minimum wins (diag < left < up on ties).  Every cell depends only on its
// three predecessors, so an anti-diagonal wavefront produces bit-identical cm / directions.
//
// Mapping: one warp per cost matrix.  Lane L (1..31) owns matrix row row0+L-1 of the current
// 31-row strip and walks it left to right; at step s it sits on column j = s-(L-1), i.e. the
// warp sweeps anti-diagonals.  Lane 0 plays the row above the strip (+inf for the first strip,
// the previous strip's last row otherwise), so `up` is one shfl_up and `diag` is last step's
// `up`; +inf stands for dtw-python's NaN "no predecessor" (never wins a strict '<').
// Local costs are staged through shared memory: each 32-column tile is read from HBM with
// fully coalesced 128-byte row segments 16..32 steps before it is consumed and parked in a
// 64-slot circular row buffer, SKEWED by the row index: element (row L, column j) lives in slot
// (j+L-1)&63 = s&63, so the read of step s is `row_base + 4*(s&63)` for every lane (an immediate
// offset, no address arithmetic in the dependent chain) and the odd row pitch of 65 words puts
// the 32 lanes on 32 distinct banks.  Directions are packed 2 bits/cell in "skewed" words (field = step index),
// one coalesced 128-byte store per 16 steps.  The backtrack then needs one step per TOKEN ROW
// (not per path cell): find the previous non-horizontal move with a clz on the packed words.
#include "common.cuh"

namespace wts {

constexpr int RS = 31;         // matrix rows per strip (lanes 1..31)
constexpr int DTW_WARPS = 4;   // warps (= matrices) per CTA
constexpr int PITCH = 65;      // shared-memory words per lane row (64 slots + 1 pad: odd pitch)
constexpr int TILE_WORDS = 33 * 64;           // per-warp staging buffer (32 rows x 65), multiple of 64 elements
#ifndef DTW_MIN_CTAS
#define DTW_MIN_CTAS 5
#endif

__host__ __device__ inline int dtw_nstrips(int T) { return (T + RS - 1) / RS; }
__host__ __device__ inline int dtw_niter(int F) { return (F + RS - 1 + 31) / 32; }
__host__ __device__ inline int dtw_wpr(int F) { return 2 * dtw_niter(F); }   // dir words per lane row

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename TIn> __device__ __forceinline__ void sts(uint32_t addr, TIn v);
template <> __device__ __forceinline__ void sts<float>(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
template <> __device__ __forceinline__ void sts<double>(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
template <typename TIn> __device__ __forceinline__ TIn lds(uint32_t addr);
template <> __device__ __forceinline__ float lds<float>(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
template <> __device__ __forceinline__ double lds<double>(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v; }

// One 31-row strip.  `tile_a` is the shared-memory byte address of this warp's staging buffer
// (aligned to 64*sizeof(TIn) so the slot index can be OR-ed in).
template <typename TIn, bool FIRST, bool WRITE_BND>
__device__ __forceinline__ void dtw_fill_strip(const TIn* __restrict__ C, const int F, const int row0,
                                               const int Ts, const int niter, const uint32_t tile_a,
                                               uint32_t* __restrict__ dirs_strip,
                                               double* __restrict__ bnd, const int lane)
{
    constexpr uint32_t ES = sizeof(TIn);             // element size
    constexpr uint32_t ROWB = PITCH * ES;            // row pitch in bytes
    constexpr uint32_t SLOTMASK = 63u * ES;
    const double INF = dinf();
    double cur = INF, upprev = INF;
    if (FIRST && lane == 1) upprev = 0.0;            // seeds cm[0,0] = 0 + lm[0,0]
    if (!FIRST && lane == 0) cur = __ldcg(bnd);      // cm[row0-1, 0]
    uint32_t acc = 0;
    const uint32_t myrow_a = tile_a + lane * ROWB;   // this lane's row
    const uint32_t laneb = lane * ES;
    TIn a[16], b[16];

    // All staging loads are unconditional: rows are clamped to the strip's last row and columns
    // to F-1, so out-of-range cells see finite duplicates (their results are never consumed).
    const char* Cb = reinterpret_cast<const char*>(C + (int64_t)row0 * F);   // warp-uniform
    const uint32_t Fb = (uint32_t)F * ES;

    // prologue: tile 0 (columns 0..31): rows 0..15 go straight to smem, rows 16..31 wait in b[]
    {
        uint32_t off = (uint32_t)min(lane, F - 1) * ES;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const TIn v = *reinterpret_cast<const TIn*>(Cb + off);
            sts<TIn>((tile_a | ((laneb + (k - 1) * ES) & SLOTMASK)) + k * ROWB, v);
            off += (k >= 1 && k < Ts) ? Fb : 0u;
        }
#pragma unroll
        for (int k = 16; k < 32; ++k) {
            b[k - 16] = *reinterpret_cast<const TIn*>(Cb + off);
            off += (k < Ts) ? Fb : 0u;
        }
    }
    double bndreg = INF, bndnext = INF;
    if (!FIRST) {
        const int idx = 1 + lane;
        bndnext = idx < F ? __ldcg(bnd + idx) : INF;
    }
    __syncwarp();

    for (int t = 0; t < niter; ++t) {
        if (!FIRST) {
            bndreg = bndnext;
            const int idx = 32 * (t + 1) + 1 + lane;
            bndnext = idx < F ? __ldcg(bnd + idx) : INF;
        }
        uint32_t off = (uint32_t)min(32 * (t + 1) + lane, F - 1) * ES;   // tile-row 1 (= strip row 0)
        const uint32_t cb = (t & 1) * 32, nb = 32 - cb;
        const uint32_t rd_a = myrow_a + cb * ES;      // slot (s & 63) of step k is rd_a + k*ES
        const uint32_t wcur = laneb + (cb + 63) * ES, wnext = laneb + (nb + 63) * ES;   // (+63 == -1 mod 64)
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            // ---- software-pipelined prefetch of tile t+1 (loads issued 16 steps before the store)
            if (k < 16) {
                sts<TIn>((tile_a | ((wcur + (16 + k) * ES) & SLOTMASK)) + (16 + k) * ROWB, b[k]);
                a[k] = *reinterpret_cast<const TIn*>(Cb + off);
            } else {
                sts<TIn>((tile_a | ((wnext + (k - 16) * ES) & SLOTMASK)) + (k - 16) * ROWB, a[k - 16]);
                b[k - 16] = *reinterpret_cast<const TIn*>(Cb + off);
            }
            off += (k >= 1 && k < Ts) ? Fb : 0u;
            __syncwarp();
            // ---- one anti-diagonal
            const int s = 32 * t + k;
            const double l = (double)lds<TIn>(rd_a + k * ES);
            const double up = __shfl_up_sync(FULL_MASK, cur, 1);
            const double diag = upprev;
            upprev = up;
            const double c1 = diag + l, c2 = cur + l, c3 = up + l;
            const bool p2 = c2 < c1;                  // left beats diag
            const double m = p2 ? c2 : c1;
            const bool p3 = c3 < m;                   // up beats both
            const double best = p3 ? c3 : m;
            cur = best;
            if ((k & 15) == 0) acc = 0;
            if (p2) acc |= 1u << (2 * (k & 15));
            if (p3) acc |= 2u << (2 * (k & 15));
            if ((k & 15) == 15) dirs_strip[(2 * t + (k >> 4)) * 32 + lane] = acc;
            if (WRITE_BND) {
                const int j = s - (Ts - 1);
                if (lane == Ts && j >= 0 && j < F) bnd[j] = best;
            }
            if (!FIRST) {
                const double ub = __shfl_sync(FULL_MASK, bndreg, k);
                if (lane == 0) cur = ub;             // cm[row0-1, s+1]
            }
        }
    }
}

// Direction fields (2 bits per cell, field index = step & 15): bit0 = "left beats diag",
// bit1 = "up beats both".  up if bit1, else left if bit0, else diag.
__device__ __forceinline__ uint32_t dtw_nonleft_mask(uint32_t x)
{
    const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
    return 0x55555555u & ~(lo & ~hi);
}

// dir field of cell (i, j)
__device__ __forceinline__ uint32_t dtw_dir_at(const uint32_t* dirs, int W, int i, int j)
{
    const int strip = i / RS, ln = i - strip * RS + 1;
    const int s = j + ln - 1;
    const uint32_t x = __ldcg(dirs + ((int64_t)strip * W + (s >> 4)) * 32 + ln);
    const uint32_t f = (x >> (2 * (s & 15))) & 3u;
    return (f & 2u) ? 3u : ((f & 1u) ? 2u : 1u);
}

template <typename TIn>
__global__ void __launch_bounds__(DTW_WARPS * 32, DTW_MIN_CTAS)
dtw_warp_kernel(const TIn* __restrict__ cost, const WtsSegDesc* __restrict__ segs, const int nseg,
                uint32_t* __restrict__ dir_ws, double* __restrict__ bnd_ws,
                int32_t* __restrict__ jumps_out, int32_t* __restrict__ path_out,
                const int64_t* __restrict__ path_off, int32_t* __restrict__ path_len)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TIn* tile = reinterpret_cast<TIn*>(smem_raw) + warp * TILE_WORDS;
    const uint32_t tile_a = smem_u32(tile);
    const int seg = blockIdx.x * DTW_WARPS + warp;
    if (seg >= nseg) return;

    const WtsSegDesc sd = segs[seg];
    const int T = sd.T, F = sd.F;
    const TIn* C = cost + sd.cost_off;
    uint32_t* dirs = dir_ws + sd.dir_off;
    double* bnd = bnd_ws + sd.bnd_off;
    int32_t* jumps = jumps_out + sd.jumps_off;
    if (T <= 0 || F <= 0) return;

    // zero the staging buffer once: cells read before their tile arrives (j < 0) must be finite
    for (int k = lane; k < 32 * PITCH; k += 32) tile[k] = TIn(0);
    __syncwarp();

    const int niter = dtw_niter(F);
    const int W = 2 * niter;
    const int nstrips = dtw_nstrips(T);
    for (int strip = 0; strip < nstrips; ++strip) {
        const int row0 = strip * RS;
        const int Ts = min(RS, T - row0);
        uint32_t* ds = dirs + (int64_t)strip * W * 32;
        const bool more = strip + 1 < nstrips;
        if (strip == 0) {
            if (more) dtw_fill_strip<TIn, true, true>(C, F, row0, Ts, niter, tile_a, ds, bnd, lane);
            else      dtw_fill_strip<TIn, true, false>(C, F, row0, Ts, niter, tile_a, ds, bnd, lane);
        } else {
            if (more) dtw_fill_strip<TIn, false, true>(C, F, row0, Ts, niter, tile_a, ds, bnd, lane);
            else      dtw_fill_strip<TIn, false, false>(C, F, row0, Ts, niter, tile_a, ds, bnd, lane);
        }
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();

    // ---- backtrack, one step per token row (executed redundantly by all lanes; lane 0 stores)
    {
        int i = T - 1, j = F - 1;
        if (lane == 0) jumps[T] = F - 1;
        while (i > 0) {
            const int strip = i / RS, ln = i - strip * RS + 1;
            const uint32_t* base = dirs + (int64_t)strip * W * 32 + ln;
            int s = j + ln - 1;
            int w = s >> 4, pos = s & 15, kf = 0;
            uint32_t x = 0;
            while (true) {
                x = __ldcg(base + w * 32);
                const uint32_t m = dtw_nonleft_mask(x) & (0xffffffffu >> (30 - 2 * pos));
                if (m) { kf = (31 - __clz(m)) >> 1; break; }
                if (w == 0) { kf = 0; break; }
                --w; pos = 15;
            }
            int jj = w * 16 + kf - (ln - 1);
            if (jj < 0) jj = 0;
            const bool is_up = (x >> (2 * kf + 1)) & 1u;
            if (lane == 0) jumps[i] = jj;
            j = (!is_up && jj > 0) ? jj - 1 : jj;       // diag or up into the previous row
            --i;
        }
        if (lane == 0) jumps[0] = 0;
    }

    // ---- optional full path (alignment.index1s / index2s), cell by cell; tests & plots only
    if (path_out != nullptr && lane == 0) {
        int32_t* p1 = path_out + path_off[seg];
        int32_t* p2 = p1 + T + F;
        int i = T - 1, j = F - 1, len = 1;
        while (i > 0 || j > 0) {
            uint32_t d = dtw_dir_at(dirs, W, i, j);
            if (i == 0) d = 2u; else if (j == 0) d = 3u;
            if (d == 1u) { --i; --j; } else if (d == 2u) { --j; } else { --i; }
            ++len;
        }
        path_len[seg] = len;
        i = T - 1; j = F - 1;
        int k = len - 1;
        p1[k] = i; p2[k] = j;
        while (i > 0 || j > 0) {
            uint32_t d = dtw_dir_at(dirs, W, i, j);
            if (i == 0) d = 2u; else if (j == 0) d = 3u;
            if (d == 1u) { --i; --j; } else if (d == 2u) { --j; } else { --i; }
            --k;
            p1[k] = i; p2[k] = j;
        }
    }
}

// status: 1 when the segment's local-cost matrix holds a non-finite value (the situation in which
// the reference's dtw() can end with "No warping path found").
template <typename TIn>
__global__ void dtw_status_kernel(const TIn* __restrict__ cost, const WtsSegDesc* __restrict__ segs,
                                  const int nseg, int32_t* __restrict__ status)
{
    const int seg = blockIdx.x;
    if (seg >= nseg) return;
    const WtsSegDesc sd = segs[seg];
    const TIn* C = cost + sd.cost_off;
    const int64_t n = (int64_t)sd.T * sd.F;
    int bad = 0;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) bad |= !isfinite((double)C[k]);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) status[seg] = bad ? 1 : 0;
}

}  // namespace wts

using namespace wts;

extern "C" int64_t wts_dtw_dir_words(int32_t T, int32_t F)
{
    if (T <= 0 || F <= 0) return 0;
    return (int64_t)dtw_nstrips(T) * dtw_wpr(F) * 32;
}

extern "C" int64_t wts_dtw_bnd_doubles(int32_t T, int32_t F)
{
    if (T <= RS || F <= 0) return 0;
    return ((int64_t)F + 3) & ~3LL;
}

extern "C" int wts_dtw_batch(const void* d_cost, int32_t cost_is_f64, const WtsSegDesc* d_segs,
                             int32_t nseg, uint32_t* d_dir_ws, double* d_bnd_ws, int32_t* d_jumps,
                             int32_t* d_path, const int64_t* d_path_off, int32_t* d_path_len,
                             int32_t* d_status, void* stream)
{
    if (nseg <= 0) return 0;
    if (!d_cost || !d_segs || !d_dir_ws || !d_jumps) { set_error("wts_dtw_batch: null pointer"); return -2; }
    if (d_path && (!d_path_off || !d_path_len)) { set_error("wts_dtw_batch: d_path needs d_path_off and d_path_len"); return -2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (nseg + DTW_WARPS - 1) / DTW_WARPS;
    if (cost_is_f64) {
        const size_t smem = (size_t)DTW_WARPS * TILE_WORDS * sizeof(double);
        WTS_CUDA_CHECK(cudaFuncSetAttribute(dtw_warp_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dtw_warp_kernel<double><<<grid, DTW_WARPS * 32, smem, st>>>(
            (const double*)d_cost, d_segs, nseg, d_dir_ws, d_bnd_ws, d_jumps, d_path, d_path_off, d_path_len);
        WTS_LAUNCH_CHECK();
        if (d_status) { dtw_status_kernel<double><<<nseg, 128, 0, st>>>((const double*)d_cost, d_segs, nseg, d_status); WTS_LAUNCH_CHECK(); }
    } else {
        const size_t smem = (size_t)DTW_WARPS * TILE_WORDS * sizeof(float);
        dtw_warp_kernel<float><<<grid, DTW_WARPS * 32, smem, st>>>(
            (const float*)d_cost, d_segs, nseg, d_dir_ws, d_bnd_ws, d_jumps, d_path, d_path_off, d_path_len);
        WTS_LAUNCH_CHECK();
        if (d_status) { dtw_status_kernel<float><<<nseg, 128, 0, st>>>((const float*)d_cost, d_segs, nseg, d_status); WTS_LAUNCH_CHECK(); }
    }
    return 0;
}
