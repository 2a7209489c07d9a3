// ffb_common.cuh -- shared device helpers and internal launcher prototypes (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define FFB_MAX_CONV 3
#define FFB_NLAYER 5

enum { FFB_ACT_TANH = 0, FFB_ACT_SWISH = 1, FFB_ACT_NONE = 2 };

namespace ffb {

// ---- scalar maths, same formulas as the reference (src/util.h:331-339):
//      logisticfv(x) = 1 / (1 + exp(-x)),  tanhfv(x) = 2 * logisticfv(2x) - 1.
// Accurate expf and IEEE division on purpose (no fast-math): the recurrence runs for
// thousands of steps and the parity tolerance is 1e-4.
__device__ __forceinline__ float logisticf(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_ref(float x) {
    const float y = logisticf(x + x);
    return (y + y) - 1.0f;
}
__device__ __forceinline__ float activate(float x, int act) {
    if (act == FFB_ACT_TANH) return tanh_ref(x);
    if (act == FFB_ACT_SWISH) return x * logisticf(x);
    return x;
}
// The tensor core TRUNCATES on every accumulate into TMEM (tools/probe_acc.py: -0.6e-7 relative per MMA on a growing sum).
// Unlike round-to-nearest noise this is a bias: it has one sign, so an integrating LSTM cell (forget gate ~ 1) adds it up
// step after step.  An accumulator that took n full-magnitude accumulations is scaled back by (1 + FFB_ACC_COMP) when it is
// read out; calibrated in profiles/r02_acc_comp.txt.
#ifndef FFB_ACC_COMP
#define FFB_ACC_COMP 0.0f
#endif
__device__ __forceinline__ float acc_comp(float v) { return FFB_ACC_COMP != 0.0f ? fmaf(v, FFB_ACC_COMP, v) : v; }

// MUFU versions for the throughput kernels: ex2.approx (2 ulp) + rcp.approx (1 ulp); same formulas as the
// reference (src/util.h:331-339), errors of a few 1e-7, far below the parity tolerances (1e-5 .. 1e-4)
__device__ __forceinline__ float fast_logistic(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
// A/B of the recurrent gates' accuracy (tools/gpu_call5.sh): the same two MUFU ops with the argument scaling and the
// reciprocal repaired in fp32 -- x * log2(e) carried as product + rounding error (folded in as 2^t * (1 + err ln 2)) and one
// Newton step on the reciprocal.  ~7 extra ALU ops per logistic, no extra MUFU.
__device__ __forceinline__ float mid_logistic(float x) {
    const float t = x * -1.4426950408889634f;
    const float err = fmaf(x, -1.4426950408889634f, -t) + x * -1.925963033500097e-8f;     // log2(e) = hi + lo
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    e = fmaf(e, err * 0.6931471805599453f, e);
    const float d = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(r, fmaf(-d, r, 1.0f), r);
}
__device__ __forceinline__ float mid_tanh(float x) {
    const float y = mid_logistic(x + x);
    return (y + y) - 1.0f;
}
__device__ __forceinline__ float fast_tanh(float x) {
    const float y = fast_logistic(x + x);
    return (y + y) - 1.0f;
}
__device__ __forceinline__ float fast_activate(float x, int act) {
    if (act == FFB_ACT_TANH) return fast_tanh(x);
    if (act == FFB_ACT_SWISH) return x * fast_logistic(x);
    return x;
}
// reference src/util.h:276-278
// run-length head row transforms (layers.c:1334-1346; softplusf util.h:83-85)
__device__ __forceinline__ float softplus_ref(float x) { return log1pf(expf(-fabsf(x))) + ((x >= 0.0f) ? x : 0.f); }
__device__ __forceinline__ float rle_head(float x, int row, float temperature) {
    if (row < 4) return 1.0f + softplus_ref(x);
    if (row < 8) return 1e-8f + softplus_ref(x);
    return 5.0f * tanhf(x) / temperature;
}
__device__ __forceinline__ float logsumexpf_ref(float x, float y) {
    return fmaxf(x, y) + log1pf(expf(-fabsf(x - y)));
}

// Tail description of the reference convolution for one (T, winlen, stride): columns
// >= tail_col0 are NOT the zero-padded textbook window; each gets up to two explicit
// terms (x_start, tap_lo, ntap), ntap == 0 meaning "no term" (bias only).
#define FFB_CONV_TAIL 32
struct ConvTail {
    int32_t tail_col0;
    int32_t x_start[FFB_CONV_TAIL][2];
    int32_t tap_lo[FFB_CONV_TAIL][2];
    int32_t ntap[FFB_CONV_TAIL][2];
};

// Per-read geometry on the device.
struct ReadGeom {
    int64_t in_off;    // first input column of this read in the layer input
    int64_t out_off;   // first output column
    int32_t T_in;
    int32_t T_out;
    int32_t tail_id;   // index into the ConvTail table
    int32_t pad;
    int64_t plane_off; // first output column in the fp16 hi/lo planes (== out_off except for the input of a tensor-core convolution)
};

}  // namespace ffb

// ---- launchers (defined in the .cu files; all asynchronous on `st`) ----------------
// Every launcher returns the number of kernels it launched (>=0) or -1 on launch error.

// conv.cu: x [cols][nf] -> y [cols'][nfilter], weights Wt [winlen*nf][nfilter].
// y (fp32) and/or the fp16 hi/lo planes yhi/ylo (x = hi + lo, for the tensor GEMM) may be NULL
// fix_cols > 0: only the first and the last fix_cols columns of every read are computed (the columns a tensor-core
// convolution over the concatenated batch gets wrong: windows reaching into a neighbouring read, and the reference's
// edge plan)
int ffb_launch_conv(const float *x, float *y, void *yhi, void *ylo, const float *Wt, const float *bias,
                    const ffb::ReadGeom *geom, const ffb::ConvTail *tails, int n_reads, int64_t total_out_cols,
                    int max_T_out, int nf, int nfilter, int winlen, int stride, int act, int fix_cols, cudaStream_t st);

// gemm.cu: C[M][N] = A[M][K] * Wt[K][N] + bias[N]   (fp32 CUDA cores)
int ffb_launch_sgemm_bias(const float *A, const float *Wt, const float *bias, float *C, int64_t M, int N, int K,
                          cudaStream_t st);
// gemm.cu: flip-flop output layer: C[M][N] = tanh(A*Wt + b) / scale  (N = 40 / 60; scale = temperature / 5)
// head = 1: run-length head instead (globalnorm_runlengthV2, layers.c:1326-1346): rows [0,4) 1 + softplus, [4,8) 1e-8 + softplus,
// the rest 5 tanhf(x) / scale with scale = temperature
int ffb_launch_ff_tanh(const float *A, const float *Wt, const float *bias, float *C, int64_t M, int N, int K,
                       float scale, int head, cudaStream_t st);

// One work item of the streamed input GEMM: a tile (ffb_gemm_tc_stream_tile_rows(K) blocks) and what it waits
// for -- the tile may be loaded once progress[idx[d]] >= cnt[d] for every d with idx[d] >= 0.  Items are
// sorted by the step at which the producing recurrent layer completes them.
struct GemmWork {
    int32_t tile;
    int32_t idx[3];
    int32_t cnt[3];
    int32_t pad;
};
#define FFB_RNN_PUBLISH_PERIOD 16   // the recurrent kernel publishes a group's progress every 16 steps

// gemm_tc.cu: tcgen05 path.  Planes are fp16 [rows][K]; W planes keep the reference's [out][in] order.
int ffb_gemm_tc_supported(int N, int K);
int ffb_launch_split_f16(const float *x, void *hi, void *lo, int64_t n, cudaStream_t st);
// n0: first column computed (0 = all of C[M][N]); n0 > 0 needs the W-stationary kernels (N % 128 == n0 % 128 == 0, K <= 384)
int ffb_launch_gemm_tc(const void *Ahi, const void *Alo, const void *Whi, const void *Wlo, const float *bias, float *C,
                       int64_t M, int N, int K, int n0, cudaStream_t st);
// Output layer on the same kernel family: trans[M][n_out] = tanh(A*W^T + b) / scale; W planes [FFB_FF_TC_ROWS][K] and
// bias [FFB_FF_TC_ROWS] are zero-padded beyond n_out (n_out = 40 / 60)
#define FFB_FF_TC_ROWS 64
int ffb_ff_tc_supported(int n_out, int K);
int ffb_launch_ff_tanh_tc(const void *Ahi, const void *Alo, const void *Whi, const void *Wlo, const float *bias, float *C,
                          int64_t M, int n_out, int K, float scale, int head, cudaStream_t st);
// Strided convolution on the tensor cores: A = the im2col view of fp16 hi/lo planes (row r = K elements from element
// r * hop), W planes [N][K] zero-padded in K, output = swish(. + b) as planes [M][N] (+ fp32 C if non-NULL)
int ffb_conv_tc_supported(int N, int K);
int ffb_launch_conv_gemm_tc(const void *Xhi, const void *Xlo, int64_t hop, const void *Whi, const void *Wlo, const float *bias,
                            float *C, void *Chi, void *Clo, int64_t M, int N, int K, cudaStream_t st);
// Streamed variant: launched (programmatic dependent launch) right behind the recurrent kernel that is still
// writing the A planes; work items are taken from per-panel ticket queues in `work` order and each waits for its dependencies;
// CTAs that find no free SM while the recurrence runs start when it ends and drain what is left.  Returns 0 (nothing launched) when the shape is unsupported.
int ffb_gemm_tc_stream_supported(int N, int K);
int ffb_gemm_tc_stream_tile_rows(int K);   // rows (blocks) per streamed tile of a projection with inner dimension K
// progress: counters published by the recurrent kernel; queue: N/128 zeroed ticket counters (one per weight panel)
int ffb_launch_gemm_tc_streamed(const void *Ahi, const void *Alo, const void *Whi, const void *Wlo, const float *bias, float *C,
                                int64_t M, int N, int K, const GemmWork *work, const int *progress, int *queue, int n0,
                                cudaStream_t st);

// rnn.cu: one recurrent layer over a ragged batch.
struct RnnBatch {
    const int32_t *order;     // [n_slots] read index per slot (sorted by length), -1 = empty
    const int64_t *blk_off;   // [n_reads+1]
    int n_slots;              // multiple of reads-per-cluster
    int n_reads;
};
int ffb_rnn_supported(int kind, int S);
int ffb_rnn_reads_per_cluster(int kind, int S);
size_t ffb_rnn_packed_floats(int kind, int S);
// pack sW [G*S][S] (row per output, reference column) into the per-CTA resident layout
void ffb_rnn_pack_weights(int kind, int S, const float *sW, float *packed);
int ffb_launch_rnn(int kind, int S, const float *Xin, const float *sW_packed, float *Hout, const RnnBatch &rb,
                   int backward, cudaStream_t st);
int ffb_rnn_prepare(int kind, int S);   // one-time function attribute setup; returns 0 or error

// rnn_tc.cu: tcgen05 recurrent layer (GRU / LSTM, S=256 at present); R = reads per cluster, multiple of 16
int ffb_rnn_tc_supported(int kind, int S);
size_t ffb_rnn_tc_image_halfs(int kind, int S);
// fused_z: NULL, or [S][S] rows of the NEXT layer's input projection (GRU: its z gate) that ride in the free quarter of the
// M=128 tile; the kernel then also writes that layer's Xin[.][0..S) (ffb_launch_rnn_tc: bnext / xnext, next_rows = 0).  For the
// top layer the rows are the flip-flop output layer's (FF_W, zero beyond nparam): next_rows = nparam, xnext = trans
// [blocks][nparam] = tanh(. + bnext) / ff_scale
void ffb_rnn_tc_pack(int kind, int S, const float *sW, uint16_t *img, const float *fused_z);
int ffb_rnn_tc_can_fuse_z(int kind, int S);
int ffb_rnn_tc_prepare(int kind, int S);
int ffb_rnn_tc_max_clusters(int kind, int S, int R);
int ffb_rnn_tc_rmax(int kind, int S);
int ffb_rnn_tc_cluster_size(int kind, int S);
size_t ffb_rnn_tc_ring_bytes(int kind, int S, int n_clusters, int R);   // L2-resident state-exchange ring
// Schedule of the tensor recurrent kernel: rb.order holds GROUPS of 16 reads (group k = order[16k .. 16k+15], longest
// first); cluster c, slot g (g < R/16) runs the groups slot_list[slot_off[c*G+g] .. slot_off[c*G+g+1]) one after the other.
struct RnnTcSched {
    const int32_t *slot_off;    // [n_clusters * G + 1]
    const int32_t *slot_list;   // group ids
    int n_clusters;
    int n_groups;
};
// progress (optional): one counter per GROUP, +1 from each of the 32 gate warps of the cluster every
// FFB_RNN_PUBLISH_PERIOD steps (and at the group's last step) once their fp16 output planes are globally visible;
// progress[n_groups] counts CTAs that have finished the layer
int ffb_launch_rnn_tc(int kind, int S, const float *Xin, const void *Wimg, float *Hout, void *Hhi, void *Hlo,
                      const RnnBatch &rb, const RnnTcSched &sched, int R, int backward, void *ring, int *progress,
                      const float *bnext, float *xnext, int next_rows, float ff_scale, cudaStream_t st);

// signal.cu: trimming + normalisation of raw reads on the device (reference src/flappie.c:251-259)
#define FFB_MAX_VARSEG_CHUNK 1024
int ffb_launch_chunk_mad(const float *raw, const int64_t *raw_off, const int64_t *chunk_off, int n_reads, int chunk,
                         int64_t total_chunks, float *mad, cudaStream_t st);
int ffb_launch_trim_bounds(const float *mad, const int64_t *raw_off, const int64_t *chunk_off, int n_reads, int chunk,
                           float perc, int64_t trim_start, int64_t trim_end, int64_t *bounds, cudaStream_t st);
int ffb_launch_normalise(const float *raw, const int64_t *raw_off, const int64_t *bounds, const int64_t *sig_off,
                         int n_reads, float delta, float *out, cudaStream_t st);

// rle.cu: the run-length ("runnie") CRF head (nr = 40)
int ffb_launch_rle_logz(float *param, const int64_t *blk_off, int n_reads, int nr, double *logZ, cudaStream_t st);   // scan + subtract
int ffb_launch_rle_viterbi(const float *param, const int64_t *blk_off, int n_reads, int nr, uint32_t *tb_scratch,
                           int32_t *path, float *qpath, float *score, cudaStream_t st);
int ffb_launch_rle_transpost(const float *param, const int64_t *blk_off, int n_reads, int nr, float *fwd_scratch,
                             float *post, cudaStream_t st);
int ffb_launch_rle_pack(const float *param, float *out, int64_t total_blocks, cudaStream_t st);

// decode.cu
int ffb_launch_logz(const float *trans, const int64_t *blk_off, int n_reads, int nr, double *logZ, cudaStream_t st);
int ffb_launch_sub_logz(float *trans, const int64_t *blk_off, int n_reads, int nr, const double *logZ,
                        int64_t total_blocks, cudaStream_t st);
int ffb_launch_viterbi(const float *trans, const int64_t *blk_off, int n_reads, int nr, uint64_t *tb_scratch,
                       int32_t *path, float *qpath, float *score, cudaStream_t st);
// fwd_scratch: 2 * (total_blocks + n_reads) * nstate floats
int ffb_launch_transpost(const float *trans, const int64_t *blk_off, int n_reads, int nr, float *fwd_scratch,
                         float *tpost, int64_t total_blocks, cudaStream_t st);
int ffb_launch_trace(const float *tpost, const int64_t *blk_off, int n_reads, int nr, void *trace, int is_log, int wide,
                     cudaStream_t st);
int ffb_launch_exp_inplace(float *x, int64_t n, cudaStream_t st);
// emit.cu: bases + quality characters per read (segments at blk_off[n] + n, NUL-terminated) and their counts; thr = the
// ascending table of qpath values at which the quality character steps up (built by the host with its own libm)
int ffb_launch_emit(const int32_t *path, const float *qpath, const int64_t *blk_off, int n_reads, int nbase, int reverse,
                    const float *thr, int nthr, char *bases, char *quals, int32_t *nbases, cudaStream_t st);
