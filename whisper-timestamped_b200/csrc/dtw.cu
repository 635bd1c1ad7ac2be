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
// Local costs are staged through shared memory: each 32-column tile is fetched from HBM with
// fully coalesced 128-byte row segments by cp.async (LDGSTS) 31 steps before it is consumed and
// parked in a 64-slot circular row buffer, SKEWED by the row index: element (row L, column j)
// lives in slot (j+L-1)&63 = s&63, so the read of step s is `row_base + 4*(s&63)` for every lane
// (an immediate offset, no address arithmetic in the dependent chain) and the odd row pitch of 65
// words puts the 32 lanes on 32 distinct banks.  Directions are packed 2 bits/cell in "skewed" words (field = step index),
// one coalesced 128-byte store per 16 steps.  The backtrack then needs one step per TOKEN ROW
// (not per path cell): find the previous non-horizontal move with a clz on the packed words.
#include "common.cuh"

namespace wts {

constexpr int RS = 31;         // matrix rows per strip (lanes 1..31)
constexpr int DTW_WARPS = 2;   // warps (= matrices) per CTA
constexpr int DS_WORDS = 24;   // direction words per lane row kept in shared memory (covers F <= 354 single-strip)
constexpr int RING = 64;       // slots per lane row: 2 tiles of 32 columns
constexpr int PITCH = RING + 1; // odd row pitch (words): the 32 lanes of a diagonal read hit 32 banks
constexpr int TILE_WORDS = 33 * 64;           // per-warp staging buffer (32 rows x 65 <= 33*64), 256-B multiple
#ifndef DTW_MIN_CTAS
#define DTW_MIN_CTAS 9
#endif

__host__ __device__ inline int dtw_nstrips(int T) { return (T + RS - 1) / RS; }
__host__ __device__ inline int dtw_niter(int F) { return (F + RS - 1 + 31) / 32; }
__host__ __device__ inline int dtw_wpr(int F) { return 2 * dtw_niter(F); }   // dir words per lane row

// segments the single-strip fast path (dtw_small_kernel, below) takes over from the general kernel
__host__ __device__ inline bool dtw_small_eligible(const WtsSegDesc& sd)
{
    return (sd.flags & (WTS_SEG_NONPOSITIVE | WTS_SEG_PITCH16)) == (WTS_SEG_NONPOSITIVE | WTS_SEG_PITCH16) &&
           sd.T >= 1 && sd.T <= RS && sd.F >= 1 && dtw_wpr(sd.F) <= DS_WORDS;
}

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename TIn> __device__ __forceinline__ TIn lds(uint32_t addr);
template <> __device__ __forceinline__ float lds<float>(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
template <> __device__ __forceinline__ double lds<double>(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v; }
template <int BYTES> __device__ __forceinline__ void cp_async(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 31-row strip.  Staging: row k of the next 32-column tile is fetched with one warp-wide cp.async (LDGSTS)
// at step k of the current tile, i.e. 31 steps before its first use, into a 64-slot ring per row, skewed by the
// row index: element (row L, column j) lives in slot (j+L-1)&63 = step&63, so step s reads
// `row_base + (s&63)*ES` for every lane (immediate offsets; the odd row pitch gives 32 distinct banks).  The slot a
// row overwrites was last read two steps earlier.  Two commit groups per tile (rows 1-15, rows 16-31) keep every
// wait_group at least 16 steps behind its loads.  Loads are clamped (column to F-1, rows beyond the strip skipped)
// so every issued address is valid; cells outside the matrix only see finite values and are never consumed.
//
// NEG: all local costs are <= 0 and cost[0,0] < 0 (what the attention post-processing produces), hence every
// accumulated cost is a strictly negative double (or +inf for "no predecessor") and  a < b  <=>  bits(a) >u bits(b).
// The two fp64 compares of the dependent chain become integer compares (same results bit for bit).
template <typename TIn, bool FIRST, bool WRITE_BND, bool NEG>
__device__ __forceinline__ void dtw_fill_strip(const TIn* __restrict__ C, const int F, const int P, const int row0,
                                               const int Ts, const int niter, const uint32_t tile_a,
                                               uint32_t* __restrict__ dirs_strip,
                                               double* __restrict__ bnd, const int lane)
{
    constexpr uint32_t ES = sizeof(TIn);
    constexpr uint32_t ROWB = PITCH * ES;
    constexpr uint32_t SLOTMASK = (RING - 1) * ES;
    const double INF = dinf();
    double cur = INF, upprev = INF;
    if (FIRST && lane == 1) upprev = 0.0;            // seeds cm[0,0] = 0 + lm[0,0]
    if (!FIRST && lane == 0) cur = __ldcg(bnd);      // cm[row0-1, 0]
    uint32_t acc = 0;
    const uint32_t myrow_a = tile_a + lane * ROWB;
    const uint32_t laneb = lane * ES;
    const char* Cb = reinterpret_cast<const char*>(C + (int64_t)row0 * P);   // strip row 0 (= tile row 1); P = row pitch
    const uint32_t Fb = (uint32_t)P * ES;

    // prologue: tile 0 (all rows), one group
    {
        const char* src = Cb + (uint32_t)min(lane, F - 1) * ES;
        for (int k = 1; k <= Ts; ++k) {
            cp_async<ES>(tile_a + k * ROWB + ((laneb + (uint32_t)(k - 1) * ES) & SLOTMASK), src);
            src += Fb;
        }
        cp_async_commit();
    }
    double bndreg = INF, bndnext = INF;
    if (!FIRST) {
        const int idx = 1 + lane;
        bndnext = idx < F ? __ldcg(bnd + idx) : INF;
    }

    for (int t = 0; t < niter; ++t) {
        if (!FIRST) {
            bndreg = bndnext;
            const int idx = 32 * (t + 1) + 1 + lane;
            bndnext = idx < F ? __ldcg(bnd + idx) : INF;
        }
        // tile t: rows 1..15 were committed >= 16 steps ago (group A), rows 16..31 are group B (waited at k = 15)
        if (t == 0) cp_async_wait<0>(); else cp_async_wait<1>();
        __syncwarp();
        const uint32_t rd_a = myrow_a + (uint32_t)((32 * t) & (RING - 1)) * ES;
        const uint32_t slot0 = laneb + (uint32_t)(32 * (t + 1) + RING - 1) * ES;   // (+RING-1 == -1 mod RING)
        const char* src = Cb + (uint32_t)min(32 * (t + 1) + lane, F - 1) * ES;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k >= 1 && k <= Ts) {                  // row k of tile t+1; its slot was last read at step k-2
                cp_async<ES>(tile_a + k * ROWB + ((slot0 + k * ES) & SLOTMASK), src);
                src += Fb;
            }
            if (k == 15) { cp_async_commit(); cp_async_wait<1>(); __syncwarp(); }   // group A out; tile t group B landed
            // ---- one anti-diagonal
            const int s = 32 * t + k;
            const double l = (double)lds<TIn>(rd_a + k * ES);
            const double up = __shfl_up_sync(FULL_MASK, cur, 1);
            const double diag = upprev;
            upprev = up;
            const double c1 = diag + l, c2 = cur + l, c3 = up + l;
            bool p2, p3;
            double m, best;
            if (NEG) {
                p2 = (unsigned long long)__double_as_longlong(c2) > (unsigned long long)__double_as_longlong(c1);
                m = p2 ? c2 : c1;
                p3 = (unsigned long long)__double_as_longlong(c3) > (unsigned long long)__double_as_longlong(m);
                best = p3 ? c3 : m;
            } else {
                p2 = c2 < c1;                         // left beats diag
                m = p2 ? c2 : c1;
                p3 = c3 < m;                          // up beats both
                best = p3 ? c3 : m;
            }
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
        cp_async_commit();                            // group B of tile t+1
    }
    cp_async_wait<0>();
}

// Direction fields (2 bits per cell, field index = step & 15): bit0 = "left beats diag",
// bit1 = "up beats both".  up if bit1, else left if bit0, else diag.
__device__ __forceinline__ uint32_t dtw_nonleft_mask(uint32_t x)
{
    const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
    return 0x55555555u & ~(lo & ~hi);
}

// ---- backtrack, one step per token row (executed redundantly by all lanes; lane 0 stores): jumps[i] = first path
// column of token row i (T.py:1648-1652), jumps[T] = F - 1.
__device__ __forceinline__ void dtw_backtrack_jumps(const uint32_t* dirs, const int W, const int T, const int F,
                                                    int32_t* __restrict__ jumps, const int lane)
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
            x = base[w * 32];
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

// dir field of cell (i, j)
__device__ __forceinline__ uint32_t dtw_dir_at(const uint32_t* dirs, int W, int i, int j)
{
    const int strip = i / RS, ln = i - strip * RS + 1;
    const int s = j + ln - 1;
    const uint32_t x = dirs[((int64_t)strip * W + (s >> 4)) * 32 + ln];
    const uint32_t f = (x >> (2 * (s & 15))) & 3u;
    return (f & 2u) ? 3u : ((f & 1u) ? 2u : 1u);
}

template <typename TIn>
__global__ void __launch_bounds__(DTW_WARPS * 32, DTW_MIN_CTAS)
dtw_warp_kernel(const TIn* __restrict__ cost, const WtsSegDesc* __restrict__ segs, const int nseg,
                uint32_t* __restrict__ dir_ws, double* __restrict__ bnd_ws,
                int32_t* __restrict__ jumps_out, int32_t* __restrict__ path_out,
                const int64_t* __restrict__ path_off, int32_t* __restrict__ path_len, const int skip_small)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TIn* tile = reinterpret_cast<TIn*>(smem_raw) + warp * TILE_WORDS;
    const uint32_t tile_a = smem_u32(tile);
    uint32_t* dirs_sm = reinterpret_cast<uint32_t*>(smem_raw + (size_t)DTW_WARPS * TILE_WORDS * sizeof(TIn)) + warp * (DS_WORDS * 32);
    const int seg = blockIdx.x * DTW_WARPS + warp;
    if (seg >= nseg) return;

    const WtsSegDesc sd = segs[seg];
    if (skip_small && dtw_small_eligible(sd) && sd.T + 1 <= (skip_small >> 8) && dtw_wpr(sd.F) <= (skip_small & 255))
        return;                                              // owned by dtw_small_kernel (skip_small = rows << 8 | dir words)
    const int T = sd.T, F = sd.F, P = seg_pitch(sd);
    const TIn* C = cost + sd.cost_off;
    // directions live in shared memory when the whole matrix fits one strip and DS_WORDS words per row
    // (the typical alignment problem): the row-wise backtrack then never waits on L2
    uint32_t* dirs = (sd.T <= RS && dtw_wpr(sd.F) <= DS_WORDS) ? dirs_sm : dir_ws + sd.dir_off;
    double* bnd = bnd_ws + sd.bnd_off;
    int32_t* jumps = jumps_out + sd.jumps_off;
    if (T <= 0 || F <= 0) return;

    // zero the staging buffer once: cells read before their tile arrives (j < 0) must be finite
    for (int k = lane; k < TILE_WORDS; k += 32) tile[k] = TIn(0);
    __syncwarp();

    const int niter = dtw_niter(F);
    const int W = 2 * niter;
    const int nstrips = dtw_nstrips(T);
    for (int strip = 0; strip < nstrips; ++strip) {
        const int row0 = strip * RS;
        const int Ts = min(RS, T - row0);
        uint32_t* ds = dirs + (int64_t)strip * W * 32;
        const bool more = strip + 1 < nstrips;
        const bool neg = (sizeof(TIn) == 4) && (sd.flags & 1);
        if (neg) {
            if (strip == 0) {
                if (more) dtw_fill_strip<TIn, true, true, true>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
                else      dtw_fill_strip<TIn, true, false, true>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
            } else {
                if (more) dtw_fill_strip<TIn, false, true, true>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
                else      dtw_fill_strip<TIn, false, false, true>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
            }
        } else if (strip == 0) {
            if (more) dtw_fill_strip<TIn, true, true, false>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
            else      dtw_fill_strip<TIn, true, false, false>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
        } else {
            if (more) dtw_fill_strip<TIn, false, true, false>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
            else      dtw_fill_strip<TIn, false, false, false>(C, F, P, row0, Ts, niter, tile_a, ds, bnd, lane);
        }
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();

    dtw_backtrack_jumps(dirs, W, T, F, jumps, lane);

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

// ------------------------------------------------------------------------------------ single-strip fast path
// The typical alignment problem (T <= 31 tokens, F <= 354 frames, float32 costs <= 0 from wts_attn_prep_batch with rows
// padded to 16 bytes) gets its own kernel, built to spend as few issue slots per anti-diagonal as the bit-exact fp64
// recurrence allows (the general kernel above is ISSUE-bound: ~40 warp instructions per step of which the recurrence
// needs ~17, profiles/r1c_dtw_summary.md):
//  * staging is 16-byte cp.async (LDGSTS.128): one warp instruction moves 8 rows x 16 columns of a tile, i.e. four
//    instructions per tile of up to 31 rows, one tile ahead, completion by cp.async groups (rows are padded to 16 bytes
//    so every chunk is aligned);
//  * the row buffers are NOT skewed: a ring of five tiles (t-2 .. t being read — the 31 lanes of skew span two
//    tiles back —, t+1 and t+2 in flight) plus a mirror of slot 0 behind the ring, so lane L reads
//    `row_base + 4 p_L + 4 k` at step k of a tile — an immediate offset with no wrap inside the tile
//    (p_L = (16 t - L + 1) mod 80 advances once per tile); row pitch 96 words: bank (p_L + k) mod 32 =
//    (1 - L + k + 16 t) mod 32 is distinct over the lanes; 15 KB of shared memory per warp (rows + directions)
//    keeps 14 warps per SM resident — the dependent fp64 chain needs them;
//  * directions stay in shared memory (2 bits per cell) for the row-wise backtrack.
// Same recurrence, same tie-breaks, same packed direction words as dtw_fill_strip<float, true, false, true>.
// Geometry of a variant <TC, LA>: TC columns per tile (= wavefront steps per iteration), LA tiles in flight ahead of the
// one being consumed.  The ring holds the tiles still being read (the current one and ceil(31 / TC) behind it: the 31
// lanes of skew), plus the LA in flight; a mirror of slot 0 sits behind the ring.  The ring length is a multiple of 32
// words, which makes bank(L) = (pitch L + 1 - L + k + TC t) mod 32 injective over the lanes for the pitches used.
template <int TC, int LA> struct SmGeo {
    static constexpr int NT = LA + 1 + (31 + TC - 1) / TC;
    static constexpr int NB = LA + 1;                        // mbarriers
    static constexpr int RING = NT * TC;                     // words
    static constexpr int PITCH = RING + TC;                  // words
    static constexpr int TILE_BYTES = 32 * PITCH * 4;        // rows 0..31 (row 0 = the virtual row above, stays zero)
    static_assert(RING % 32 == 0, "ring must be a multiple of 32 words (bank-conflict-free reads)");
    static_assert(16 % TC == 0 || TC % 16 == 0, "tile must divide or be a multiple of a direction word");
    static constexpr int warp_bytes(bool dirs_smem) { return TILE_BYTES + (dirs_smem ? DS_WORDS * 32 * 4 : 0) + 64; }
};

// float32 -> float64 widening on the INTEGER pipe (exact for zeros and normal numbers; `ok` = false for denormals, inf and
// nan, which the caller widens with the fp64-pipe conversion instead).  Experiment WTS_DTW_VARIANT=5: takes one of the four
// fp64-pipe instructions of an anti-diagonal step (3 DADD + 1 F2F) off that pipe.
__device__ __forceinline__ double widen_f32_int(uint32_t u, bool& ok)
{
    const uint32_t mag = u & 0x7fffffffu;
    const uint32_t e = mag >> 23;
    ok = (e != 255u) && (e != 0u || mag == 0u);
    const uint32_t hi = (mag == 0u ? 0u : (mag >> 3) + 0x38000000u) | (u & 0x80000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}

template <int TC, int LA, bool DIRS_SMEM, bool INTW = false>
__global__ void __launch_bounds__(32)
dtw_small_kernel(const float* __restrict__ cost, const WtsSegDesc* __restrict__ segs, const int nseg,
                 uint32_t* __restrict__ dir_ws, int32_t* __restrict__ jumps_out, const int l2_prefetch, const int n_rows,
                 const int n_dir_words)
{
    // n_rows: row buffers in shared memory (1 + the largest T of the batch, <= 32); n_dir_words: direction words per lane
    // (for the largest F of the batch, <= DS_WORDS) — sized per launch so that more warps fit an SM
    using G = SmGeo<TC, LA>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int seg = blockIdx.x;                              // one warp (= one CTA) per matrix
    if (seg >= nseg) return;
    const WtsSegDesc sd = segs[seg];
    if (!dtw_small_eligible(sd) || sd.T + 1 > n_rows || dtw_wpr(sd.F) > n_dir_words) return;   // the general kernel owns it
    unsigned char* my = smem_raw;
    const uint32_t tile_a = smem_u32(my);
    const int tile_bytes = n_rows * G::PITCH * 4;
    uint32_t* dirs = DIRS_SMEM ? reinterpret_cast<uint32_t*>(my + tile_bytes) : dir_ws + sd.dir_off;

    const int T = sd.T, F = sd.F, P = (F + 3) & ~3;
    const float* C = cost + sd.cost_off;
    // Optional (WTS_DTW_L2PF=1): one bulk L2 prefetch of the whole contiguous matrix (T x P float32, <= 44 KB), so that
    // the 64..128-byte tile copies hit L2.  Measured: no effect (0.329 vs 0.323 ms for 16384 matrices) — the kernel is
    // not bound by the DRAM access pattern.
    if (lane == 0 && l2_prefetch)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(C), "r"((uint32_t)(T * P * 4)) : "memory");
    // zero the row buffers once: cells read before their tile arrives (j < 0), rows beyond T and the virtual row 0
    // must hold finite values (INF + 0 stays INF; they never feed a cell of the matrix)
    {
        float4* z = reinterpret_cast<float4*>(my);
        for (int k = lane; k < tile_bytes / 16; k += 32) z[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    const int niter = dtw_niter(F);                          // 32-step tiles of the wavefront (direction layout: 2 words each)
    const int nit = niter * (32 / TC);                       // TC-step iterations
    const int ntile = (P + TC - 1) / TC;                     // column tiles of the matrix
    // Staging: 16-byte cp.async (LDGSTS.128).  A lane owns chunk `cq` (4 columns) of the rows rq, rq + RPI, ... of every
    // tile: one warp instruction moves 32 chunks = RPI rows x TC columns, so a tile of T <= 31 rows costs
    // ceil(T / RPI) instructions (twice that when its ring slot is mirrored) plus their address arithmetic.
    // (A per-lane cp.async.bulk looked like "one instruction for all rows" in PTX, but SASS serialises it: ELECT + R2UR +
    // UBLKCP + branch per active lane, ~10 issue slots per copy — 17 of the 39 instructions per step in the ncu capture.)
    constexpr int CPR = TC / 4;                              // 16-byte chunks per row and tile
    constexpr int RPI = 32 / CPR;                            // rows per warp instruction
    constexpr int NG = (31 + RPI - 1) / RPI;                 // instructions per tile
    const int cq = lane % CPR, rq = lane / CPR;
    const float* src_lane = C + (int64_t)rq * P + 4 * cq;
    const uint32_t dst_lane = tile_a + (uint32_t)(rq + 1) * (G::PITCH * 4) + 16u * cq;

    auto issue_tile = [&](int u) {                           // columns [TC u, TC u + TC) of every row -> ring slot u % NT
        if (u < ntile) {
            const int slot = u % G::NT;
            const bool col_ok = TC * u + 4 * cq < P;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (col_ok && rq + RPI * g < T) {
                    const float* src = src_lane + (int64_t)(RPI * g) * P + TC * u;
                    const uint32_t dst = dst_lane + (uint32_t)(RPI * g) * (G::PITCH * 4) + (uint32_t)(slot * TC * 4);
                    cp_async<16>(dst, src);
                    if (slot == 0) cp_async<16>(dst + G::RING * 4, src);      // mirror behind the ring
                }
            }
        }
        cp_async_commit();                                   // one group per tile, also when nothing was issued
    };

    const double INF = dinf();
    double cur = INF, upprev = INF;
    if (lane == 1) upprev = 0.0;                             // seeds cm[0,0] = 0 + lm[0,0]
    uint32_t pb = (uint32_t)((G::RING - lane + 1) % G::RING) * 4u;   // 4 * ((TC t - L + 1) mod RING), t = 0
    uint32_t acc = 0;
    const uint32_t myrow_a = tile_a + min(lane, n_rows - 1) * (G::PITCH * 4);   // lanes beyond the buffers re-read the last row
#pragma unroll
    for (int u = 0; u < LA; ++u) issue_tile(u);
    for (int t = 0; t < nit; ++t) {
        cp_async_wait<LA - 1>();                             // all but the newest LA - 1 groups have landed: tile t is in
        __syncwarp();
        issue_tile(t + LA);                                  // its ring slot held a tile last read in iteration t - 1
        const uint32_t rd = myrow_a + pb;
        const int s0 = (t * TC) & 15;                        // position of this tile inside its direction word
        if (s0 == 0) acc = 0;
#pragma unroll
        for (int k = 0; k < TC; ++k) {
            double l;
            if (INTW) {
                const float lf = lds<float>(rd + 4 * k);
                bool ok;
                l = widen_f32_int(__float_as_uint(lf), ok);
                if (__any_sync(FULL_MASK, !ok)) l = (double)lf;      // denormal / non-finite somewhere in the warp: rare
            } else {
                l = (double)lds<float>(rd + 4 * k);
            }
            const double up = __shfl_up_sync(FULL_MASK, cur, 1);
            const double c1 = upprev + l, c2 = cur + l, c3 = up + l;
            upprev = up;
            // all sums are strictly negative doubles or +inf: a < b  <=>  bits(a) >u bits(b)
            const bool p2 = (unsigned long long)__double_as_longlong(c2) > (unsigned long long)__double_as_longlong(c1);
            const double m = p2 ? c2 : c1;
            const bool p3 = (unsigned long long)__double_as_longlong(c3) > (unsigned long long)__double_as_longlong(m);
            cur = p3 ? c3 : m;
            if (TC >= 16) {
                if ((k & 15) == 0) acc = 0;
                if (p2) acc |= 1u << (2 * (k & 15));
                if (p3) acc |= 2u << (2 * (k & 15));
                if ((k & 15) == 15) dirs[((t * TC + k) >> 4) * 32 + lane] = acc;
            } else {
                if (p2) acc |= (1u << (2 * k)) << (2 * s0);
                if (p3) acc |= (2u << (2 * k)) << (2 * s0);
            }
        }
        if (TC < 16 && s0 + TC == 16) dirs[((t * TC) >> 4) * 32 + lane] = acc;
        pb += TC * 4;
        if (pb >= (uint32_t)(G::RING * 4)) pb -= G::RING * 4;
    }
    cp_async_wait<0>();
    if (!DIRS_SMEM) __threadfence_block();
    __syncwarp();
    dtw_backtrack_jumps(dirs, 2 * niter, T, F, jumps_out + sd.jumps_off, lane);
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
    const int64_t n = (int64_t)sd.T * seg_pitch(sd);       // padding columns hold zeros (finite)
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

extern "C" int wts_dtw_batch_sized(const void* d_cost, int32_t cost_is_f64, const WtsSegDesc* d_segs,
                                   int32_t nseg, uint32_t* d_dir_ws, double* d_bnd_ws, int32_t* d_jumps,
                                   int32_t* d_path, const int64_t* d_path_off, int32_t* d_path_len,
                                   int32_t* d_status, int32_t max_T, int32_t max_F, void* stream);

extern "C" int wts_dtw_batch(const void* d_cost, int32_t cost_is_f64, const WtsSegDesc* d_segs,
                             int32_t nseg, uint32_t* d_dir_ws, double* d_bnd_ws, int32_t* d_jumps,
                             int32_t* d_path, const int64_t* d_path_off, int32_t* d_path_len,
                             int32_t* d_status, void* stream)
{
    return wts_dtw_batch_sized(d_cost, cost_is_f64, d_segs, nseg, d_dir_ws, d_bnd_ws, d_jumps, d_path, d_path_off, d_path_len,
                               d_status, 0, 0, stream);
}

extern "C" int wts_dtw_batch_sized(const void* d_cost, int32_t cost_is_f64, const WtsSegDesc* d_segs,
                                   int32_t nseg, uint32_t* d_dir_ws, double* d_bnd_ws, int32_t* d_jumps,
                                   int32_t* d_path, const int64_t* d_path_off, int32_t* d_path_len,
                                   int32_t* d_status, int32_t max_T, int32_t max_F, void* stream)
{
    // max_T / max_F: largest T / F of the batch when the caller knows them (0 = unknown): they size the shared-memory
    // buffers of the single-strip fast path (fewer rows / direction words -> more resident warps)
    const int n_rows = (max_T > 0 && max_T < RS) ? max_T + 1 : RS + 1;
    const int n_dir_words = (max_F > 0 && dtw_wpr(max_F) < DS_WORDS) ? dtw_wpr(max_F) : DS_WORDS;
    if (nseg <= 0) return 0;
    if (!d_cost || !d_segs || !d_dir_ws || !d_jumps) { set_error("wts_dtw_batch: null pointer"); return -2; }
    if (d_path && (!d_path_off || !d_path_len)) { set_error("wts_dtw_batch: d_path needs d_path_off and d_path_len"); return -2; }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (nseg + DTW_WARPS - 1) / DTW_WARPS;
    if (cost_is_f64) {
        const size_t smem = (size_t)DTW_WARPS * (TILE_WORDS * sizeof(double) + DS_WORDS * 32 * sizeof(uint32_t));
        WTS_CUDA_CHECK(cudaFuncSetAttribute(dtw_warp_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dtw_warp_kernel<double><<<grid, DTW_WARPS * 32, smem, st>>>(
            (const double*)d_cost, d_segs, nseg, d_dir_ws, d_bnd_ws, d_jumps, d_path, d_path_off, d_path_len, 0);
        WTS_LAUNCH_CHECK();
        if (d_status) { dtw_status_kernel<double><<<nseg, 128, 0, st>>>((const double*)d_cost, d_segs, nseg, d_status); WTS_LAUNCH_CHECK(); }
    } else {
        // single-strip fast path (dtw_small_kernel) for the segments that qualify; it produces jumps only, so a
        // request for full paths keeps everything in the general kernel.  WTS_DTW_SMALL=0 turns it off.
        static const int small_on = [] { const char* e = getenv("WTS_DTW_SMALL"); return e ? atoi(e) : 1; }();
        const int use_small = small_on && d_path == nullptr;
        const size_t smem = (size_t)DTW_WARPS * (TILE_WORDS * sizeof(float) + DS_WORDS * 32 * sizeof(uint32_t));
        WTS_CUDA_CHECK(cudaFuncSetAttribute(dtw_warp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dtw_warp_kernel<float><<<grid, DTW_WARPS * 32, smem, st>>>(
            (const float*)d_cost, d_segs, nseg, d_dir_ws, d_bnd_ws, d_jumps, d_path, d_path_off, d_path_len,
            use_small ? ((n_rows << 8) | n_dir_words) : 0);
        WTS_LAUNCH_CHECK();
        if (use_small) {
            // geometry variants (WTS_DTW_VARIANT, default 0): <columns per tile, tiles in flight, directions in shared memory>
            static const int variant = [] { const char* e = getenv("WTS_DTW_VARIANT"); return e ? atoi(e) : 4; }();
            static const int l2pf = [] { const char* e = getenv("WTS_DTW_L2PF"); return e ? atoi(e) : 0; }();   // measured: no effect
#define WTS_LAUNCH_SMALL(TC_, LA_, DS_, IW_)                                                                              \
            do {                                                                                                          \
                const size_t smem_s = (size_t)n_rows * SmGeo<TC_, LA_>::PITCH * 4 + (DS_ ? n_dir_words * 32 * 4 : 0) + 64;      \
                WTS_CUDA_CHECK(cudaFuncSetAttribute(dtw_small_kernel<TC_, LA_, DS_, IW_>,                                 \
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));          \
                dtw_small_kernel<TC_, LA_, DS_, IW_><<<nseg, 32, smem_s, st>>>((const float*)d_cost, d_segs, nseg,       \
                                                                              d_dir_ws, d_jumps, l2pf, n_rows,           \
                                                                              n_dir_words);                              \
            } while (0)
            switch (variant) {
                case 1: WTS_LAUNCH_SMALL(16, 1, false, false); break;
                case 2: WTS_LAUNCH_SMALL(8, 3, true, false); break;
                case 3: WTS_LAUNCH_SMALL(8, 3, false, false); break;
                case 4: WTS_LAUNCH_SMALL(32, 1, true, false); break;
                case 5: WTS_LAUNCH_SMALL(16, 1, false, true); break;
                case 0: WTS_LAUNCH_SMALL(16, 1, true, false); break;
                default: WTS_LAUNCH_SMALL(32, 1, true, false); break;        // measured best
            }
#undef WTS_LAUNCH_SMALL
            WTS_LAUNCH_CHECK();
        }
        if (d_status) { dtw_status_kernel<float><<<nseg, 128, 0, st>>>((const float*)d_cost, d_segs, nseg, d_status); WTS_LAUNCH_CHECK(); }
    }
    return 0;
}
