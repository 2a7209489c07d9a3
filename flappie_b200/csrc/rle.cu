// rle.cu -- the run-length ("runnie") CRF head over a ragged batch.
//
// Replaces reference
//   runlengthV2_partition_function + the "-= logZ" of globalnorm_runlengthV2   (src/layers.c:1255-1358)
//   decode_crf_runlength                                                       (src/decode.c:901-984)
//   transpost_crf_runlength                                                    (src/decode.c:1013-1159)
// (the affine map + softplus / tanh row transforms of globalnorm_runlengthV2 are the ACT = 2 epilogue of
// gemm_tc_kernel / the `rle` mode of ff_tanh_kernel).
//
// Layout: param / post are [block][nr] row-major, nr = 2 nbase + 2 nbase^2 (40): rows [0, nbase) shape,
// [nbase, 2 nbase) scale, then the transition scores at  2 nbase + to * 2 nbase + from + (stay_from ? nbase : 0)
// (rle_trans_lookup, decode.c:893-898).  States: b < nbase "move into base b", b + nbase "stay in base b"; a move
// state may go to any OTHER base or into its own stay state, a stay state likewise.
//
// One warp per read, strictly sequential in the block index, the next block's scores prefetched into registers.
// The Viterbi visits candidates in the reference's order (b2 ascending, move before stay, strict '>'), so the path
// and the score are bit-exact on bit-identical input; the posteriors fold their logsumexp terms in the reference's
// order with the reference's association of the three-term sums.
#include "ffb_common.cuh"

namespace ffb {

constexpr unsigned RFULL = 0xffffffffu;
constexpr int RB = 4;            // bases
constexpr int RS = 8;            // states
constexpr int RNR = 40;          // rows per block

__device__ __forceinline__ int rle_idx(int from_base, int stay_from, int to_base) {
    return RS + to_base * RS + from_base + (stay_from ? RB : 0);      // offset within a block's 40 rows
}
// the j-th base other than b (j = 0..2), ascending
__device__ __forceinline__ int other_base(int b, int j) { return j + (j >= b ? 1 : 0); }

__device__ __forceinline__ double shfl_d(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(RFULL, lo, src); hi = __shfl_sync(RFULL, hi, src);
    return __hiloint2double(hi, lo);
}

// ---------------------------------------------------------------------------------
// log partition function in DOUBLE (layers.c:1255-1302).  Lane = b1 * 8 + k: k < 6 is the k-th term of move state b1
// (source base other_base(b1, k / 2), from its move (k even) or stay (k odd) state); the stay states go through the
// FLOAT logsumexpf, as in the reference.  Every lane keeps the whole state vector.
__global__ void __launch_bounds__(32)
rle_logz_kernel(const float *__restrict__ param, const int64_t *__restrict__ blk_off, int n_reads, double *__restrict__ logZ) {
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) { if (lane == 0) logZ[rd] = 0.0; return; }
    const float *p = param + b0 * RNR;
    const int b1 = lane >> 3, k = lane & 7;
    const bool term = k < 6;
    const int b2 = other_base(b1, (k >> 1) % 3), st = k & 1;
    const int my_idx = rle_idx(b2, st, b1), my_src = b2 + st * RB;
    double prev[RS];
#pragma unroll
    for (int s = 0; s < RS; s++) prev[s] = 0.0;             // calloc, layers.c:1262
    float pm = term ? p[my_idx] : 0.0f, ps[RS];
#pragma unroll
    for (int s = 0; s < RS; s++) ps[s] = p[rle_idx(s % RB, s / RB, s % RB)];      // the eight same-base scores
    for (int c = 0; c < T; c++) {
        const float cm = pm;
        float cs[RS];
#pragma unroll
        for (int s = 0; s < RS; s++) cs[s] = ps[s];
        if (c + 1 < T) {
            const float *pn = p + (int64_t)(c + 1) * RNR;
            pm = term ? pn[my_idx] : 0.0f;
#pragma unroll
            for (int s = 0; s < RS; s++) ps[s] = pn[rle_idx(s % RB, s / RB, s % RB)];
        }
        double src = 0.0;
#pragma unroll
        for (int s = 0; s < RS; s++) if (s == my_src) src = prev[s];
        const double v = term ? src + (double)cm : -HUGE_VAL;
        double m = v;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmax(m, shfl_d(m, lane ^ o));
        double e = term ? exp(v - m) : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) e += shfl_d(e, lane ^ o);
        const double mv = m + log(e);                        // new curr[b1], in every lane of the segment
        double curr[RS];
#pragma unroll
        for (int b = 0; b < RB; b++) {
            curr[b] = shfl_d(mv, b * 8);
            const float x = (float)(prev[b] + (double)cs[b]);              // move -> own stay
            const float y = (float)(prev[b + RB] + (double)cs[b + RB]);    // stay -> stay
            curr[b + RB] = (double)logsumexpf_ref(x, y);
        }
#pragma unroll
        for (int s = 0; s < RS; s++) prev[s] = curr[s];
    }
    if (lane == 0) {
        double z = prev[0];
        for (int s = 1; s < RS; s++) z = fmax(z, prev[s]) + log1p(exp(-fabs(z - prev[s])));
        logZ[rd] = z;
    }
}

// param[blk][r] -= (float)(logZ / (float)T) for the transition rows r >= 2 nbase (layers.c:1349-1356)
__global__ void rle_sub_logz_kernel(float *__restrict__ param, const int64_t *__restrict__ blk_off, int n_reads,
                                    const double *__restrict__ logZ) {
    const int rd = blockIdx.y;
    if (rd >= n_reads) return;
    const int64_t b0 = blk_off[rd];
    const int64_t T = blk_off[rd + 1] - b0;
    if (T <= 0) return;
    const float lz = (float)(logZ[rd] / (float)T);
    const int64_t n = T * (RNR - RS);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t blk = i / (RNR - RS);
        const int r = RS + (int)(i % (RNR - RS));
        param[(b0 + blk) * RNR + r] -= lz;
    }
}

// ---------------------------------------------------------------------------------
// Viterbi (decode.c:901-984).  Lane = b1 * 8 + k as above; candidates of a move state are visited in k order with a
// strict '>' in the reference, i.e. the lowest k among equal maxima wins.  Traceback: 4 bits per state, one 32-bit
// word per block.  path[T] (the layout keeps T+1 slots per read) is set to -1.
__global__ void __launch_bounds__(32)
rle_viterbi_kernel(const float *__restrict__ param, const int64_t *__restrict__ blk_off, int n_reads, uint32_t *__restrict__ tb_scratch,
                   int32_t *__restrict__ path, float *__restrict__ qpath, float *__restrict__ score) {
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) { if (lane == 0) score[rd] = NAN; return; }
    const float *p = param + b0 * RNR;
    uint32_t *tb = tb_scratch + b0;
    const int b1 = lane >> 3, k = lane & 7;
    const bool term = k < 6;
    const int b2 = other_base(b1, (k >> 1) % 3), st = k & 1;
    const int my_idx = rle_idx(b2, st, b1), my_src = b2 + st * RB;
    float prev[RS];
#pragma unroll
    for (int s = 0; s < RS; s++) prev[s] = 0.0f;            // calloc, decode.c:909
    float pm = term ? p[my_idx] : 0.0f, ps[RS];
#pragma unroll
    for (int s = 0; s < RS; s++) ps[s] = p[rle_idx(s % RB, s / RB, s % RB)];
    for (int c = 0; c < T; c++) {
        const float cm = pm;
        float cs[RS];
#pragma unroll
        for (int s = 0; s < RS; s++) cs[s] = ps[s];
        if (c + 1 < T) {
            const float *pn = p + (int64_t)(c + 1) * RNR;
            pm = term ? pn[my_idx] : 0.0f;
#pragma unroll
            for (int s = 0; s < RS; s++) ps[s] = pn[rle_idx(s % RB, s / RB, s % RB)];
        }
        float src = 0.0f;
#pragma unroll
        for (int s = 0; s < RS; s++) if (s == my_src) src = prev[s];
        float v = term ? src + cm : -INFINITY;
        int kk = k;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(RFULL, v, o);
            const int ok = __shfl_xor_sync(RFULL, kk, o);
            if (ov > v || (ov == v && ok < kk)) { v = ov; kk = ok; }
        }
        // (v, kk): best score and its candidate for move state b1, in every lane of the segment
        const int from_mv = other_base(b1, (kk >> 1) % 3) + (kk & 1) * RB;
        float curr[RS];
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < RB; b++) {
            curr[b] = __shfl_sync(RFULL, v, b * 8);
            word |= (uint32_t)__shfl_sync(RFULL, from_mv, b * 8) << (4 * b);
            const float sv = prev[b + RB] + cs[b + RB];      // stay -> stay
            const float mvv = prev[b] + cs[b];               // move -> own stay
            const bool from_stay = sv > mvv;                 // decode.c:955-963: ties come from the move state
            curr[b + RB] = from_stay ? sv : mvv;
            word |= (uint32_t)(from_stay ? b + RB : b) << (4 * (b + RB));
        }
        if (lane == 0) tb[c] = word;
#pragma unroll
        for (int s = 0; s < RS; s++) prev[s] = curr[s];
    }
    __syncwarp();
    if (lane == 0) {
        int last = 0;                                        // argmaxf: first max wins (util.c:17-31)
        for (int s = 1; s < RS; s++) if (prev[s] > prev[last]) last = s;
        score[rd] = prev[last];
        int32_t *pp = path + b0 + rd;
        float *qq = qpath + b0 + rd;
        pp[T] = -1; qq[T] = 0.0f;
        for (int blk = T; blk > 0; blk--) {
            const uint32_t w = tb[blk - 1];
            pp[blk - 1] = last;
            qq[blk - 1] = 0.0f;                              // the run-length decoder has no per-block quality
            last = (int)((w >> (4 * last)) & 15u);
        }
    }
}

// ---------------------------------------------------------------------------------
// Posteriors (decode.c:1013-1159): forward vectors, then the backward pass emits fwd + bwd + score (NOT normalised) and
// copies the shape / scale rows.  Lane = state (lanes >= 8 idle); folds in the reference's order.
__global__ void __launch_bounds__(32)
rle_fwd_kernel(const float *__restrict__ param, const int64_t *__restrict__ blk_off, int n_reads, float *__restrict__ fwd) {
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x, s = lane & 7;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) return;
    const float *p = param + b0 * RNR;
    float *rf = fwd + (b0 + rd) * RS;                         // (T+1) x 8
    const int b = s % RB;
    const bool is_stay = s >= RB;
    float P = 0.0f;                                           // make_flappie_matrix zero-fills, decode.c:1021
    if (lane < RS) rf[lane] = 0.0f;
    // this destination's scores: move state b <- (b2 stay, b2 move) for the three other bases; stay state b <- (b stay, b move)
    int idx[6];
#pragma unroll
    for (int j = 0; j < 3; j++) { idx[2 * j] = rle_idx(other_base(b, j), 1, b); idx[2 * j + 1] = rle_idx(other_base(b, j), 0, b); }
    int src[6];                                               // source state of each score
#pragma unroll
    for (int j = 0; j < 3; j++) { src[2 * j] = other_base(b, j) + RB; src[2 * j + 1] = other_base(b, j); }
    if (is_stay) { idx[0] = rle_idx(b, 1, b); idx[1] = rle_idx(b, 0, b); src[0] = b + RB; src[1] = b; }
    float pn[6];
#pragma unroll
    for (int j = 0; j < 6; j++) pn[j] = p[idx[j]];
    for (int c = 0; c < T; c++) {
        float pc[6];
#pragma unroll
        for (int j = 0; j < 6; j++) pc[j] = pn[j];
        if (c + 1 < T) {
#pragma unroll
            for (int j = 0; j < 6; j++) pn[j] = p[(int64_t)(c + 1) * RNR + idx[j]];
        }
        // all shuffles are executed by the whole warp; only the arithmetic depends on the kind of state
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; j++) v[j] = __shfl_sync(RFULL, P, src[j]) + pc[j];
        float np;
        if (is_stay) {
            np = logsumexpf_ref(v[0], v[1]);                                   // (stay, move)
        } else {
            np = -INFINITY;
#pragma unroll
            for (int j = 0; j < 3; j++) np = logsumexpf_ref(np, logsumexpf_ref(v[2 * j], v[2 * j + 1]));
        }
        P = np;
        if (lane < RS) rf[(int64_t)(c + 1) * RS + lane] = P;
    }
}

__global__ void __launch_bounds__(32)
rle_bwd_kernel(const float *__restrict__ param, const int64_t *__restrict__ blk_off, int n_reads, const float *__restrict__ fwd,
               float *__restrict__ post) {
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x, s = lane & 7;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) return;
    const float *p = param + b0 * RNR;
    float *q = post + b0 * RNR;
    const float *rf = fwd + (b0 + rd) * RS;
    const int b = s % RB, stay_from = s / RB;
    // this source state's transitions: to the three other bases (move states), then into its own base's stay state
    int idx[4];
#pragma unroll
    for (int j = 0; j < 3; j++) idx[j] = rle_idx(b, stay_from, other_base(b, j));
    idx[3] = rle_idx(b, stay_from, b);
    float B = 0.0f;                                           // calloc, decode.c:1023
    float pn[4], fn = rf[(int64_t)(T - 1) * RS + s];
#pragma unroll
    for (int j = 0; j < 4; j++) pn[j] = p[(int64_t)(T - 1) * RNR + idx[j]];
    for (int blk = T - 1; blk >= 0; blk--) {
        float pc[4];
#pragma unroll
        for (int j = 0; j < 4; j++) pc[j] = pn[j];
        const float f = fn;
        if (blk > 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) pn[j] = p[(int64_t)(blk - 1) * RNR + idx[j]];
            fn = rf[(int64_t)(blk - 1) * RS + s];
        }
        float nb = -INFINITY;
        float *qc = q + (int64_t)blk * RNR;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float Bt = __shfl_sync(RFULL, B, other_base(b, j));       // prev[b2]
            nb = logsumexpf_ref(nb, Bt + pc[j]);
            if (lane < RS) qc[idx[j]] = f + Bt + pc[j];                      // fwd + prev + param   (decode.c:1096,1100)
        }
        const float Bs = __shfl_sync(RFULL, B, b + RB);                      // prev[b + nbase]
        nb = logsumexpf_ref(nb, Bs + pc[3]);
        if (lane < RS) {
            qc[idx[3]] = f + pc[3] + Bs;                                     // fwd + param + prev   (decode.c:1108,1114)
            qc[lane] = p[(int64_t)blk * RNR + lane];                         // shape / scale rows copied through (:1119-1122)
        }
        B = nb;
    }
}

// rle_params[blk][0..7] = param[blk][0..7]: the shape and scale rows runnie prints (runnie.c:293-296)
__global__ void rle_pack_kernel(const float *__restrict__ param, float *__restrict__ out, int64_t total_blocks) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_blocks * RS; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = param[(i / RS) * RNR + (i % RS)];
}

}  // namespace ffb

#define RLE_OK(n) (cudaGetLastError() == cudaSuccess ? (n) : -1)

int ffb_launch_rle_logz(float *param, const int64_t *blk_off, int n_reads, int nr, double *logZ, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    if (nr != ffb::RNR) return -1;
    ffb::rle_logz_kernel<<<n_reads, 32, 0, st>>>(param, blk_off, n_reads, logZ);
    ffb::rle_sub_logz_kernel<<<dim3(16, n_reads), 256, 0, st>>>(param, blk_off, n_reads, logZ);
    return RLE_OK(2);
}

int ffb_launch_rle_viterbi(const float *param, const int64_t *blk_off, int n_reads, int nr, uint32_t *tb_scratch,
                           int32_t *path, float *qpath, float *score, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    if (nr != ffb::RNR) return -1;
    ffb::rle_viterbi_kernel<<<n_reads, 32, 0, st>>>(param, blk_off, n_reads, tb_scratch, path, qpath, score);
    return RLE_OK(1);
}

// fwd_scratch: (total_blocks + n_reads) * 8 floats
int ffb_launch_rle_transpost(const float *param, const int64_t *blk_off, int n_reads, int nr, float *fwd_scratch,
                             float *post, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    if (nr != ffb::RNR) return -1;
    ffb::rle_fwd_kernel<<<n_reads, 32, 0, st>>>(param, blk_off, n_reads, fwd_scratch);
    ffb::rle_bwd_kernel<<<n_reads, 32, 0, st>>>(param, blk_off, n_reads, fwd_scratch, post);
    return RLE_OK(2);
}

int ffb_launch_rle_pack(const float *param, float *out, int64_t total_blocks, cudaStream_t st) {
    if (total_blocks <= 0) return 0;
    const int64_t n = total_blocks * ffb::RS;
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    ffb::rle_pack_kernel<<<grid, 256, 0, st>>>(param, out, total_blocks);
    return RLE_OK(1);
}
