/*
 * libwts — C-ABI of the B200-native word-alignment hot path of whisper-timestamped.
 *
 * Conventions (all entry points):
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (PyTorch allocations are
 *     only the memory carrier); the library never frees or retains it past the call, except
 *     the persistent objects created by *_create and freed by *_destroy;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously on it;
 *   - return value: 0 = OK, <0 = error (message via wts_last_error(), thread-local);
 *   - no C++ exception crosses the ABI; there is no CPU fallback: if no sm_100 device code can
 *     run, the call fails with an error.
 *
 * Each entry cites the reference interface it replaces
 * (T.py = /root/reference/whisper_timestamped/transcribe.py).
 */
#ifndef WTS_H
#define WTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WTS_VERSION 100

/* One alignment problem (= one speech segment handed to perform_word_alignment, T.py:1428).
 * Built on the host, copied to the device by the caller, consumed by the kernels. */
typedef struct WtsSegDesc {
    int32_t window;    /* which decoded 30-s window of the qk buffer the rows live in            */
    int32_t row0;      /* first token row inside that window's qk rows                          */
    int32_t last_row;  /* row used for the LAST token (normally row0+T-1; differs when the text
                          is truncated because T > F, T.py:1516-1535)                           */
    int32_t T;         /* number of tokens (rows of the cost matrix)                            */
    int32_t f0;        /* start_token: first encoder frame of the slice (T.py:1540)             */
    int32_t F;         /* number of frames in the slice (= end_token - start_token)             */
    int32_t max_dur;   /* padding limit in frames (find_start_padding(mfcc)//2, T.py:1556-1558),
                          <=0 when there is no padding                                          */
    int32_t flags;     /* reserved, 0                                                           */
    int64_t cost_off;  /* element offset of this segment's [T,F] matrix in the cost buffer      */
    int64_t jumps_off; /* element offset of this segment's T+1 jumps in the jumps buffer        */
    int64_t dir_off;   /* uint32 offset of this segment's direction words in the DTW workspace  */
    int64_t bnd_off;   /* float64 offset of the strip-boundary row in the DTW workspace
                          (only read when T > 32)                                               */
} WtsSegDesc;

int         wts_version(void);
const char* wts_last_error(void);

/* Workspace sizing helpers for wts_dtw_batch (pure host arithmetic).
 * dir words (uint32) and boundary doubles one segment of T tokens x F frames needs. */
int64_t wts_dtw_dir_words(int32_t T, int32_t F);
int64_t wts_dtw_bnd_doubles(int32_t T, int32_t F);

/*
 * Fused attention post-processing — replaces T.py:1540-1568
 *   (slice frames, stack alignment heads, scipy.ndimage.median_filter(1,1,9), softmax over
 *    frames, mean over heads, / L2-norm over tokens, negate, padding mask, cost[0,0]=min).
 * d_qk:   float32 [n_windows, N, Tmax, Fmax] pre-softmax cross-attention rows of the N selected
 *         alignment heads (what hook_attention_weights captures at T.py:783-793), row k = the
 *         row computed with input token k.
 * d_cost: float32 output, matrices back to back at segs[i].cost_off (values are exactly the
 *         float32 numbers the reference widens to float64 at T.py:1550).
 * max_T / max_F: maxima of segs[i].T / segs[i].F over the batch (launch geometry only).
 */
int wts_attn_prep_batch(const float* d_qk, int32_t N, int32_t Tmax, int32_t Fmax,
                        const WtsSegDesc* d_segs, int32_t nseg, int32_t max_T, int32_t max_F,
                        float* d_cost, void* stream);

/*
 * Batched monotonic DTW — replaces dtw.dtw(weights, step_pattern=symmetric1) at T.py:1572-1581
 * and the jumps extraction at T.py:1648-1652.
 * d_cost:  float32 (cost_is_f64 == 0) or float64 (cost_is_f64 == 1) local-cost matrices.
 * d_dir_ws: uint32 workspace, segment i uses wts_dtw_dir_words(T,F) words at dir_off.
 * d_bnd_ws: float64 workspace, segment i uses wts_dtw_bnd_doubles(T,F) doubles at bnd_off.
 * d_jumps: int32 output, T+1 entries per segment at jumps_off (first frame of every token row on
 *          the optimal path, then the last frame index).
 * d_path:  optional (may be NULL) int32 output of the full warping path (alignment.index1s /
 *          .index2s): segment i writes index1s at d_path[path_off[i] .. +len) and index2s at
 *          d_path[path_off[i] + T + F .. +len), ascending; len goes to d_path_len[i].
 * d_status: optional int32 per segment, 0 = ok, 1 = non-finite accumulated cost (the reference
 *          raises "No warping path found" in that case).
 */
int wts_dtw_batch(const void* d_cost, int32_t cost_is_f64,
                  const WtsSegDesc* d_segs, int32_t nseg,
                  uint32_t* d_dir_ws, double* d_bnd_ws,
                  int32_t* d_jumps, int32_t* d_path, const int64_t* d_path_off,
                  int32_t* d_path_len, int32_t* d_status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WTS_H */
