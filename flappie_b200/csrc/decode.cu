// decode.cu -- flip-flop CRF decoding over a ragged batch: global normalisation constant,
// Viterbi, transition posteriors (forward/backward) and the state trace.
//
// Replaces reference
//   crf_manystay_partition_function + the "-= logZ" of globalnorm_manystay (src/layers.c:1035-1096)
//   decode_crf_flipflop + trans_lookup + argmaxf/valmaxf   (src/decode.c:104-204, src/util.c:17-61)
//   transpost_crf_flipflop + log_row_normalise_inplace      (src/decode.c:377-497, src/flappie_matrix.c:450-467)
//   exp_activation_inplace + trace_from_posterior           (src/layers.c:56-66, src/decode.c:499-543)
//
// Layout: trans / tpost are [block][nr] row-major (one reference column per row),
// nr = nstate*(nbase+1): rows b1*nstate+from for flip destination b1, then nstate entries
// at nbase*nstate: entry f < nbase = flip f -> flop f+nbase, entry f >= nbase = flop stay.
//
// All scans are strictly sequential in the block index (max-plus / log-sum-exp
// recurrences); parallelism is one WARP PER READ, lanes = (destination, source) pairs,
// with the read's scores streamed through shared memory by cp.async double buffering.
// fp32 additions and comparisons are done in the reference's visit order so the Viterbi
// path is bit-exact given bit-identical input (first-max-wins ties included).
#include <cstdlib>

#include "ffb_common.cuh"

namespace ffb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int DEC_CHUNK = 16;   // blocks per cp.async stage

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Stage `nblk` blocks (nr floats each, 16-byte aligned rows) of one read into smem.
__device__ __forceinline__ void stage_chunk(float *dst, const float *src, int nblk, int nr, int lane) {
    const int n16 = nblk * nr / 4;
    for (int i = lane; i < n16; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
    cp_async_commit();
}

// ---------------------------------------------------------------------------------
// Viterbi.  One warp per read.
//   NBASE == 4: single pass, lane = b1*8 + from reads trans[lane] directly.
//   otherwise : lanes = two 16-lane segments, segment h handles destination 2*it+h.
// Traceback pointers: 4 bits per state packed into one 64-bit word per block.
template <int NBASE>
__global__ void __launch_bounds__(32)
viterbi_kernel(const float *__restrict__ trans, const int64_t *__restrict__ blk_off, int n_reads,
               uint64_t *__restrict__ tb, int32_t *__restrict__ path, float *__restrict__ qpath,
               float *__restrict__ score) {
    constexpr int NSTATE = 2 * NBASE;
    constexpr int NR = NSTATE * (NBASE + 1);
    __shared__ __align__(16) float stage[2][DEC_CHUNK * NR];
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    int32_t *rpath = path + b0 + rd;       // T+1 entries
    float *rqpath = qpath + b0 + rd;
    if (T <= 0) {
        if (lane == 0) { score[rd] = nanf(""); }
        return;
    }
    const float *tr = trans + b0 * NR;
    uint64_t *rtb = tb + b0;

    // lane roles
    const int from = (NBASE == 4) ? (lane & 7) : (lane & 15);
    const int seg = (NBASE == 4) ? (lane >> 3) : (lane >> 4);
    float P = 0.0f;   // score of state `from` after the previous block (calloc, decode.c:130)

    const int nchunk = (T + DEC_CHUNK - 1) / DEC_CHUNK;
    stage_chunk(stage[0], tr, min(DEC_CHUNK, T), NR, lane);
    uint64_t tbword = 0;   // lane i keeps the word of block (32*k + i)
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = ch * DEC_CHUNK;
        const int cn = min(DEC_CHUNK, T - c0);
        if (ch + 1 < nchunk) {
            stage_chunk(stage[(ch + 1) & 1], tr + (int64_t)(c0 + DEC_CHUNK) * NR, min(DEC_CHUNK, T - c0 - DEC_CHUNK), NR, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float *sb = stage[ch & 1];
        for (int i = 0; i < cn; i++) {
            const float *col = sb + i * NR;
            const float *flop = col + NSTATE * NBASE;
            uint64_t nib = 0;
            float Pnew = P;
            // ---- flop destinations (decode.c:153-164): stay is the default, move wins on '>' ----
            {
                const float Pm = __shfl_sync(FULL, P, (lane - NBASE) & 31);   // prev[from - nbase] lives NBASE lanes down
                if (from >= NBASE && from < NSTATE) {
                    const float stay = P + flop[from];
                    const float move = Pm + flop[from - NBASE];
                    const bool mv = move > stay;
                    Pnew = mv ? move : stay;
                    if (seg == 0) nib |= (uint64_t)(mv ? from - NBASE : from) << (4 * from);
                }
            }
            // ---- flip destinations (decode.c:167-180): source 0 is the default, later sources win on '>' ----
            if constexpr (NBASE == 4) {
                const float v = col[lane] + P;
                float m = v;
                m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
                m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
                m = fmaxf(m, __shfl_xor_sync(FULL, m, 4));
                const unsigned eq = __ballot_sync(FULL, v == m);
                const int win = __ffs((eq >> (8 * seg)) & 0xffu) - 1;   // first max wins
                if (from == 0) nib |= (uint64_t)win << (4 * seg);
                const float mf = __shfl_sync(FULL, m, (from & 3) * 8);  // curr[from] for from < 4
                if (from < NBASE) Pnew = mf;
            } else {
                constexpr int NIT = (NBASE + 1) / 2;
#pragma unroll
                for (int it = 0; it < NIT; it++) {
                    const int b1 = 2 * it + seg;
                    const bool valid = (from < NSTATE) && (b1 < NBASE);
                    const float v = valid ? col[b1 * NSTATE + from] + P : -INFINITY;
                    float m = v;
                    m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
                    m = fmaxf(m, __shfl_xor_sync(FULL, m, 4));
                    m = fmaxf(m, __shfl_xor_sync(FULL, m, 8));
                    const unsigned eq = __ballot_sync(FULL, valid && v == m);
                    const int win = __ffs((eq >> (16 * seg)) & 0xffffu) - 1;
                    if (from == 0 && b1 < NBASE) nib |= (uint64_t)win << (4 * b1);
                    // states 2*it and 2*it+1 were just produced by segments 0 and 1
                    const float m0 = __shfl_sync(FULL, m, 0);
                    const float m1 = __shfl_sync(FULL, m, 16);
                    if (from == 2 * it) Pnew = m0;
                    if (from == 2 * it + 1 && from < NBASE) Pnew = m1;
                }
            }
            P = Pnew;
            // gather the 4-bit pointers of all states into one word
            const unsigned lo = __reduce_or_sync(FULL, (unsigned)(nib & 0xffffffffu));
            const unsigned hi = (NSTATE > 8) ? __reduce_or_sync(FULL, (unsigned)(nib >> 32)) : 0u;
            const int blk = c0 + i;
            if ((blk & 31) == lane) tbword = ((uint64_t)hi << 32) | lo;
            if ((blk & 31) == 31 || blk == T - 1) {
                const int base = blk & ~31;
                if (base + lane <= blk) rtb[base + lane] = tbword;
            }
        }
        __syncwarp();
    }

    // ---- final state: valmaxf / argmaxf, first max wins (util.c:17-31, decode.c:184-185) ----
    {
        const bool holder = (seg == 0) && (from < NSTATE);
        const float v = holder ? P : -INFINITY;
        float m = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        const unsigned eq = __ballot_sync(FULL, holder && v == m);
        const int st = __ffs(eq) - 1;   // lane index == state for segment 0
        if (lane == 0) {
            score[rd] = m;
            rpath[T] = st;
        }
        // ---- traceback (decode.c:186-191), 32 blocks per round ----
        int state = st;
        __syncwarp();
        for (int hi_blk = T; hi_blk > 0; hi_blk -= 32) {
            const int lo_blk = max(hi_blk - 32, 0);   // blocks [lo_blk, hi_blk)
            const int myblk = lo_blk + lane;
            const uint64_t w = (myblk < hi_blk) ? rtb[myblk] : 0ull;
            int mine = 0;
            for (int i = hi_blk - 1 - lo_blk; i >= 0; i--) {
                const unsigned wl = __shfl_sync(FULL, (unsigned)(w & 0xffffffffu), i);
                const unsigned wh = (NSTATE > 8) ? __shfl_sync(FULL, (unsigned)(w >> 32), i) : 0u;
                const uint64_t ww = ((uint64_t)wh << 32) | wl;
                state = (int)((ww >> (4 * state)) & 0xfull);   // path[blk] = tb[blk][path[blk+1]]
                if (lane == i) mine = state;
            }
            if (myblk < hi_blk) rpath[myblk] = mine;
        }
    }
    __syncwarp();
    // ---- qpath (decode.c:190-192): score of the transition taken into block blk ----
    for (int blk = lane; blk <= T; blk += 32) {
        if (blk == 0) {
            rqpath[0] = nanf("");
        } else {
            const int pf = rpath[blk - 1], pt = rpath[blk];
            const int idx = (pt < NBASE) ? (pt * NSTATE + pf) : (NBASE * NSTATE + pf);   // trans_lookup
            rqpath[blk] = tr[(int64_t)(blk - 1) * NR + idx];
        }
    }
}

// ---------------------------------------------------------------------------------
// log partition function in DOUBLE (layers.c:1035-1079).  One warp per read; 16-lane
// segments, segment h handles flip destination 2*it+h.  The flip sums use the
// max-shifted form (equal to the reference's sequential logsumexp fold to ~1e-16
// relative, far below the float the result is rounded to, layers.c:1089).
template <int NBASE>
__global__ void __launch_bounds__(32)
logz_kernel(const float *__restrict__ C, const int64_t *__restrict__ blk_off, int n_reads, double *__restrict__ logZ) {
    constexpr int NSTATE = 2 * NBASE;
    constexpr int NR = NSTATE * (NBASE + 1);
    __shared__ __align__(16) float stage[2][DEC_CHUNK * NR];
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) {
        if (lane == 0) logZ[rd] = 0.0;
        return;
    }
    const float *tr = C + b0 * NR;
    const int from = lane & 15, seg = lane >> 4;
    double P = 0.0;   // calloc, layers.c:1041
    const int nchunk = (T + DEC_CHUNK - 1) / DEC_CHUNK;
    stage_chunk(stage[0], tr, min(DEC_CHUNK, T), NR, lane);
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = ch * DEC_CHUNK;
        const int cn = min(DEC_CHUNK, T - c0);
        if (ch + 1 < nchunk) {
            stage_chunk(stage[(ch + 1) & 1], tr + (int64_t)(c0 + DEC_CHUNK) * NR, min(DEC_CHUNK, T - c0 - DEC_CHUNK), NR, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float *sb = stage[ch & 1];
        for (int i = 0; i < cn; i++) {
            const float *col = sb + i * NR;
            const float *stay = col + NSTATE * NBASE;
            double Pnew = P;
            const double Pm = __shfl_sync(FULL, P, (lane - NBASE) & 31);
            if (from >= NBASE && from < NSTATE) {
                const double x = P + (double)stay[from];
                const double y = Pm + (double)stay[from - NBASE];
                Pnew = fmax(x, y) + log1p(exp(-fabs(x - y)));   // util.h:280-282
            }
            constexpr int NIT = (NBASE + 1) / 2;
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                const int b1 = 2 * it + seg;
                const bool valid = (from < NSTATE) && (b1 < NBASE);
                const double v = valid ? (double)col[b1 * NSTATE + from] + P : -INFINITY;
                double m = v;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
                double e = valid ? exp(v - m) : 0.0;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) e += __shfl_xor_sync(FULL, e, o);
                const double r = m + log(e);
                const double r0 = __shfl_sync(FULL, r, 0);
                const double r1 = __shfl_sync(FULL, r, 16);
                if (from == 2 * it) Pnew = r0;
                if (from == 2 * it + 1 && from < NBASE) Pnew = r1;
            }
            P = Pnew;
        }
        __syncwarp();
    }
    // logZ = logsumexp over final states (layers.c:1071-1074)
    const bool holder = (seg == 0) && (from < NSTATE);
    const double v = holder ? P : -INFINITY;
    double m = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
    double e = holder ? exp(v - m) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) e += __shfl_xor_sync(FULL, e, o);
    if (lane == 0) logZ[rd] = m + log(e);
}

// trans[blk][*] -= (float)(logZ / T)   (layers.c:1089-1096)
__global__ void sub_logz_kernel(float *__restrict__ trans, const int64_t *__restrict__ blk_off, int n_reads, int nr,
                                const double *__restrict__ logZ) {
    const int rd = blockIdx.y;
    const int64_t b0 = blk_off[rd];
    const int64_t n = (blk_off[rd + 1] - b0) * nr;
    if (n <= 0) return;
    const float lz = (float)(logZ[rd] / (double)(blk_off[rd + 1] - b0));
    float *p = trans + b0 * nr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] -= lz;
}

// ---------------------------------------------------------------------------------
// Transition posteriors, shift-invariant ("fb" kernels).
//
// tpost[blk][e] = fwd[from, blk] + bwd[to, blk+1] + trans[blk][e], log-normalised per block
// (decode.c:377-497, flappie_matrix.c:450-467).  Any constant added to a forward row, a
// backward row or a block of trans cancels in that per-block normalisation, so
//   * the scans run on running-shifted vectors (state 0 pinned to 0 every step): values stay
//     O(10), which also makes them MORE accurate than the reference's drifting fp32 sums;
//   * the global normalisation constant logZ/T (layers.c:1035-1096) is not needed for the
//     posteriors at all: in forward-backward mode the fp64 partition scan and the "-= logZ/T"
//     pass are skipped unless the caller asks for `trans` itself.
// The reference's sequential logsumexp folds become max-shifted sums (one exp per term in
// parallel, one log): same value to ~1e-7, an eighth of the dependent latency.
// One warp per read.  Lane = (destination slot lane / SEG, source lane % SEG); NBASE = 4 covers
// the 32 flip transitions in one pass, NBASE = 5 takes three passes of 2 x 16 lanes.
template <int NBASE>
struct FbGeom {
    static constexpr int NSTATE = 2 * NBASE;
    static constexpr int NR = NSTATE * (NBASE + 1);
    static constexpr int SEG = (NSTATE <= 8) ? 8 : 16;
    static constexpr int DPP = 32 / SEG;                       // flip destinations per pass
    static constexpr int NPASS = (NBASE + DPP - 1) / DPP;
};
template <int SEG>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
    for (int o = SEG / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
template <int SEG>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
    for (int o = SEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

template <int NBASE>
__global__ void __launch_bounds__(32)
fb_fwd_kernel(const float *__restrict__ trans, const int64_t *__restrict__ blk_off, int n_reads, float *__restrict__ fwd) {
    using Gm = FbGeom<NBASE>;
    constexpr int NSTATE = Gm::NSTATE, NR = Gm::NR, SEG = Gm::SEG, DPP = Gm::DPP, NPASS = Gm::NPASS;
    __shared__ __align__(16) float stage[2][DEC_CHUNK * NR];
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) return;
    const float *tr = trans + b0 * NR;
    float *rf = fwd + (b0 + rd) * NSTATE;   // (T+1) x NSTATE, row t = forward vector before block t (shifted)
    const int src = lane % SEG, slot = lane / SEG;
    const bool src_ok = src < NSTATE;
    float P = 0.0f;                          // P[src], replicated in every segment; fwd[.,0] = 0
    if (lane < NSTATE) rf[lane] = 0.0f;
    const int nchunk = (T + DEC_CHUNK - 1) / DEC_CHUNK;
    stage_chunk(stage[0], tr, min(DEC_CHUNK, T), NR, lane);
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = ch * DEC_CHUNK;
        const int cn = min(DEC_CHUNK, T - c0);
        if (ch + 1 < nchunk) {
            stage_chunk(stage[(ch + 1) & 1], tr + (int64_t)(c0 + DEC_CHUNK) * NR, min(DEC_CHUNK, T - c0 - DEC_CHUNK), NR, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float *sb = stage[ch & 1];
        for (int i = 0; i < cn; i++) {
            const float *col = sb + i * NR;
            // flop destinations src >= NBASE: logsumexp(stay, move from flip src - NBASE)   (decode.c:404-411)
            const float fl = src_ok ? col[NBASE * NSTATE + src] : 0.0f;
            const float x = P + fl;                                               // stay (valid where src >= NBASE)
            const float y = __shfl_sync(FULL, x, (lane - NBASE) & 31);            // P[src-NBASE] + flop[src-NBASE]
            const float flopv = fmaxf(x, y) + log1pf(expf(-fabsf(x - y)));
            // flip destinations: logsumexp over all sources                            (decode.c:414-422)
            float flipv[NPASS];
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const int b1 = p * DPP + slot;
                const bool ok = src_ok && b1 < NBASE;
                const float v = ok ? P + col[b1 * NSTATE + src] : -INFINITY;
                const float m = seg_max<SEG>(v);
                const float sum = seg_sum<SEG>(ok ? expf(v - m) : 0.0f);
                flipv[p] = m + logf(sum);                                         // same in every lane of the segment
            }
            // redistribute: new P[src] = flip value of destination src (src < NBASE) or the flop value
            float np = flopv;
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const float g = __shfl_sync(FULL, flipv[p], (src % DPP) * SEG);
                if (src < NBASE && src / DPP == p) np = g;
            }
            const float ref = __shfl_sync(FULL, flipv[0], 0);                     // new value of state 0: pin it to 0
            P = src_ok ? np - ref : 0.0f;
            if (lane < NSTATE) rf[(int64_t)(c0 + i + 1) * NSTATE + lane] = P;
        }
        __syncwarp();
    }
}

template <int NBASE>
__global__ void __launch_bounds__(32)
fb_bwd_kernel(const float *__restrict__ trans, const int64_t *__restrict__ blk_off, int n_reads,
              const float *__restrict__ fwd, float *__restrict__ tpost) {
    using Gm = FbGeom<NBASE>;
    constexpr int NSTATE = Gm::NSTATE, NR = Gm::NR, SEG = Gm::SEG, DPP = Gm::DPP, NPASS = Gm::NPASS;
    __shared__ __align__(16) float stage[2][DEC_CHUNK * NR];
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int lane = threadIdx.x;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) return;
    const float *tr = trans + b0 * NR;
    float *tp = tpost + b0 * NR;
    const float *rf = fwd + (b0 + rd) * NSTATE;
    const int src = lane % SEG, slot = lane / SEG;
    const bool src_ok = src < NSTATE;
    const int to_flop = src < NBASE ? src + NBASE : src;      // flop-row entry `src`: flip src -> flop src+NBASE, or flop stay
    float B = 0.0f;                                           // bwd[src] after the current block, replicated; calloc -> 0
    // chunks are walked from the end of the read; chunk `ch` covers blocks [ch*DEC_CHUNK, ...)
    const int nchunk = (T + DEC_CHUNK - 1) / DEC_CHUNK;
    {
        const int cl = nchunk - 1;
        stage_chunk(stage[cl & 1], tr + (int64_t)cl * DEC_CHUNK * NR, T - cl * DEC_CHUNK, NR, lane);
    }
    float f_next = src_ok ? rf[(int64_t)(T - 1) * NSTATE + src] : 0.0f;   // forward row of the block about to be processed
    for (int ch = nchunk - 1; ch >= 0; ch--) {
        const int c0 = ch * DEC_CHUNK;
        const int cn = min(DEC_CHUNK, T - c0);
        if (ch > 0) {
            stage_chunk(stage[(ch - 1) & 1], tr + (int64_t)(c0 - DEC_CHUNK) * NR, DEC_CHUNK, NR, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float *sb = stage[ch & 1];
        for (int i = cn - 1; i >= 0; i--) {
            const int blk = c0 + i;
            const float *col = sb + i * NR;
            const float f = f_next;
            if (blk > 0) f_next = src_ok ? rf[(int64_t)(blk - 1) * NSTATE + src] : 0.0f;   // prefetch
            // ---- terms: t = trans + bwd[to]; x = t + fwd[from] ----
            const float Bto = __shfl_sync(FULL, B, to_flop);            // (shuffles stay outside divergent selects)
            const float tfl = src_ok ? col[NBASE * NSTATE + src] + Bto : -INFINITY;   // flop row
            float t[NPASS], x[NPASS];
            float m = (slot == 0 && src_ok) ? tfl + f : -INFINITY;       // the flop row is counted once (segment 0)
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const int b1 = p * DPP + slot;
                const bool ok = src_ok && b1 < NBASE;
                const float Bb = __shfl_sync(FULL, B, b1 < NBASE ? b1 : 0);
                t[p] = ok ? col[b1 * NSTATE + src] + Bb : -INFINITY;
                x[p] = ok ? t[p] + f : -INFINITY;
                m = fmaxf(m, x[p]);
            }
            // ---- per-block log normalisation over all NR entries (flappie_matrix.c:450-467) ----
            m = seg_max<32>(m);
            float sum = (slot == 0 && src_ok) ? expf(tfl + f - m) : 0.0f;
#pragma unroll
            for (int p = 0; p < NPASS; p++) sum += (x[p] > -INFINITY) ? expf(x[p] - m) : 0.0f;
            sum = seg_sum<32>(sum);
            const float lse = m + logf(sum);
            float *pc = tp + (int64_t)blk * NR;
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const int b1 = p * DPP + slot;
                if (src_ok && b1 < NBASE) pc[b1 * NSTATE + src] = x[p] - lse;
            }
            if (slot == 0 && src_ok) pc[NBASE * NSTATE + src] = tfl + f - lse;
            // ---- backward update: bwd[from = src] = logsumexp over destinations (decode.c:465-482) ----
            float m2 = tfl;
#pragma unroll
            for (int p = 0; p < NPASS; p++) m2 = fmaxf(m2, t[p]);
#pragma unroll
            for (int o = SEG; o < 32; o <<= 1) m2 = fmaxf(m2, __shfl_xor_sync(FULL, m2, o));   // across segments: same src
            float s2 = 0.0f;
#pragma unroll
            for (int p = 0; p < NPASS; p++) s2 += (t[p] > -INFINITY) ? expf(t[p] - m2) : 0.0f;
#pragma unroll
            for (int o = SEG; o < 32; o <<= 1) s2 += __shfl_xor_sync(FULL, s2, o);
            s2 += src_ok ? expf(tfl - m2) : 0.0f;                        // the flop term, once per lane
            const float nb = src_ok ? m2 + logf(s2) : 0.0f;
            const float ref = __shfl_sync(FULL, nb, 0);                  // pin state 0 to 0
            B = src_ok ? nb - ref : 0.0f;
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------
// Production layout of the posteriors: the two scans are independent of each other, so they run CONCURRENTLY --
// one CTA of two warps per read, warp 0 the forward scan, warp 1 the backward scan, each writing its running-shifted
// state vectors -- and a third, fully block-parallel kernel (HBM-bound: 160 B read + 64 B vectors + 160 B written per
// block) combines them, tpost[blk][e] = (trans[blk][e] + bwd[to, blk+1]) + fwd[from, blk], and log-normalises each
// block.  Replaces fb_fwd_kernel -> fb_bwd_kernel (two dependent latency-bound scans, the second one also carrying
// the normalisation on its chain).
template <int NBASE>
__device__ __forceinline__ void fb_bwd_scan(const float *__restrict__ tr, int T, float *__restrict__ rb /* (T+1) x NSTATE */,
                                            float (*stage)[DEC_CHUNK * FbGeom<NBASE>::NR], int lane) {
    using Gm = FbGeom<NBASE>;
    constexpr int NSTATE = Gm::NSTATE, NR = Gm::NR, SEG = Gm::SEG, DPP = Gm::DPP, NPASS = Gm::NPASS;
    const int src = lane % SEG, slot = lane / SEG;
    const bool src_ok = src < NSTATE;
    const int to_flop = src < NBASE ? src + NBASE : src;
    float B = 0.0f;
    if (lane < NSTATE) rb[(int64_t)T * NSTATE + lane] = 0.0f;     // bwd[., T] = 0 (the reference's calloc, decode.c:440)
    const int nchunk = (T + DEC_CHUNK - 1) / DEC_CHUNK;
    {
        const int cl = nchunk - 1;
        stage_chunk(stage[cl & 1], tr + (int64_t)cl * DEC_CHUNK * NR, T - cl * DEC_CHUNK, NR, lane);
    }
    for (int ch = nchunk - 1; ch >= 0; ch--) {
        const int c0 = ch * DEC_CHUNK;
        const int cn = min(DEC_CHUNK, T - c0);
        if (ch > 0) {
            stage_chunk(stage[(ch - 1) & 1], tr + (int64_t)(c0 - DEC_CHUNK) * NR, DEC_CHUNK, NR, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float *sb = stage[ch & 1];
        for (int i = cn - 1; i >= 0; i--) {
            const float *col = sb + i * NR;
            // t = trans + bwd[to]; bwd[from = src] = logsumexp over destinations (decode.c:465-482)
            const float Bto = __shfl_sync(FULL, B, to_flop);
            const float tfl = src_ok ? col[NBASE * NSTATE + src] + Bto : -INFINITY;
            float t[NPASS];
            float m2 = tfl;
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const int b1 = p * DPP + slot;
                const bool ok = src_ok && b1 < NBASE;
                const float Bb = __shfl_sync(FULL, B, b1 < NBASE ? b1 : 0);
                t[p] = ok ? col[b1 * NSTATE + src] + Bb : -INFINITY;
                m2 = fmaxf(m2, t[p]);
            }
#pragma unroll
            for (int o = SEG; o < 32; o <<= 1) m2 = fmaxf(m2, __shfl_xor_sync(FULL, m2, o));   // across segments: same src
            float s2 = 0.0f;
#pragma unroll
            for (int p = 0; p < NPASS; p++) s2 += (t[p] > -INFINITY) ? expf(t[p] - m2) : 0.0f;
#pragma unroll
            for (int o = SEG; o < 32; o <<= 1) s2 += __shfl_xor_sync(FULL, s2, o);
            s2 += src_ok ? expf(tfl - m2) : 0.0f;                        // the flop term, once per lane
            const float nb = src_ok ? m2 + logf(s2) : 0.0f;
            const float ref = __shfl_sync(FULL, nb, 0);                  // pin state 0 to 0
            B = src_ok ? nb - ref : 0.0f;
            if (lane < NSTATE) rb[(int64_t)(c0 + i) * NSTATE + lane] = B;
        }
        __syncwarp();
    }
}

template <int NBASE>
__device__ __forceinline__ void fb_fwd_scan(const float *__restrict__ tr, int T, float *__restrict__ rf,
                                            float (*stage)[DEC_CHUNK * FbGeom<NBASE>::NR], int lane) {
    using Gm = FbGeom<NBASE>;
    constexpr int NSTATE = Gm::NSTATE, NR = Gm::NR, SEG = Gm::SEG, DPP = Gm::DPP, NPASS = Gm::NPASS;
    const int src = lane % SEG, slot = lane / SEG;
    const bool src_ok = src < NSTATE;
    float P = 0.0f;
    if (lane < NSTATE) rf[lane] = 0.0f;
    const int nchunk = (T + DEC_CHUNK - 1) / DEC_CHUNK;
    stage_chunk(stage[0], tr, min(DEC_CHUNK, T), NR, lane);
    for (int ch = 0; ch < nchunk; ch++) {
        const int c0 = ch * DEC_CHUNK;
        const int cn = min(DEC_CHUNK, T - c0);
        if (ch + 1 < nchunk) {
            stage_chunk(stage[(ch + 1) & 1], tr + (int64_t)(c0 + DEC_CHUNK) * NR, min(DEC_CHUNK, T - c0 - DEC_CHUNK), NR, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float *sb = stage[ch & 1];
        for (int i = 0; i < cn; i++) {
            const float *col = sb + i * NR;
            const float fl = src_ok ? col[NBASE * NSTATE + src] : 0.0f;
            const float x = P + fl;
            const float y = __shfl_sync(FULL, x, (lane - NBASE) & 31);
            const float flopv = fmaxf(x, y) + log1pf(expf(-fabsf(x - y)));
            float flipv[NPASS];
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const int b1 = p * DPP + slot;
                const bool ok = src_ok && b1 < NBASE;
                const float v = ok ? P + col[b1 * NSTATE + src] : -INFINITY;
                const float m = seg_max<SEG>(v);
                const float sum = seg_sum<SEG>(ok ? expf(v - m) : 0.0f);
                flipv[p] = m + logf(sum);
            }
            float np = flopv;
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const float g = __shfl_sync(FULL, flipv[p], (src % DPP) * SEG);
                if (src < NBASE && src / DPP == p) np = g;
            }
            const float ref = __shfl_sync(FULL, flipv[0], 0);
            P = src_ok ? np - ref : 0.0f;
            if (lane < NSTATE) rf[(int64_t)(c0 + i + 1) * NSTATE + lane] = P;
        }
        __syncwarp();
    }
}

// fwd / bwd: (T+1) x NSTATE rows per read starting at row blk_off[rd] + rd
template <int NBASE>
__global__ void __launch_bounds__(64)
fb_scan2_kernel(const float *__restrict__ trans, const int64_t *__restrict__ blk_off, int n_reads, float *__restrict__ fwd,
                float *__restrict__ bwd) {
    using Gm = FbGeom<NBASE>;
    __shared__ __align__(16) float stage[2][2][DEC_CHUNK * Gm::NR];
    const int rd = blockIdx.x;
    if (rd >= n_reads) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) return;
    const float *tr = trans + b0 * Gm::NR;
    if (warp == 0) fb_fwd_scan<NBASE>(tr, T, fwd + (b0 + rd) * Gm::NSTATE, stage[0], lane);
    else fb_bwd_scan<NBASE>(tr, T, bwd + (b0 + rd) * Gm::NSTATE, stage[1], lane);
}

// one warp per block: entries e = lane, lane + 32 (< NR); per-block log normalisation (flappie_matrix.c:450-467)
template <int NBASE>
__global__ void __launch_bounds__(256)
fb_combine_kernel(const float *__restrict__ trans, const int64_t *__restrict__ blk_off, int n_reads, const float *__restrict__ fwd,
                  const float *__restrict__ bwd, float *__restrict__ tpost, int64_t total_blocks) {
    using Gm = FbGeom<NBASE>;
    constexpr int NSTATE = Gm::NSTATE, NR = Gm::NR;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // (from, to) of this lane's two entries
    int from[2], to[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int e = lane + 32 * k;
        if (e < NBASE * NSTATE) { from[k] = e % NSTATE; to[k] = e / NSTATE; }
        else { const int j = e - NBASE * NSTATE; from[k] = j; to[k] = j < NBASE ? j + NBASE : j; }
    }
    // each warp walks a contiguous range of blocks: the owning read is searched once and then advances linearly
    const int64_t per = (total_blocks + nwarp - 1) / nwarp;
    const int64_t blk_lo = warp0 * per, blk_hi = (blk_lo + per < total_blocks) ? blk_lo + per : total_blocks;
    if (blk_lo >= blk_hi) return;
    int rd = 0;
    {
        int lo = 0, hi = n_reads;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (blk_off[mid] <= blk_lo) lo = mid; else hi = mid;
        }
        rd = lo;
    }
    // the block's scores are prefetched one iteration ahead (the loop is otherwise one HBM latency per block)
    float cn[2];
    cn[0] = trans[blk_lo * NR + lane];
    cn[1] = (lane + 32 < NR) ? trans[blk_lo * NR + lane + 32] : 0.0f;
    for (int64_t blk = blk_lo; blk < blk_hi; blk++) {
        while (blk >= blk_off[rd + 1]) rd++;
        const float c0v = cn[0], c1v = cn[1];
        if (blk + 1 < blk_hi) {
            cn[0] = trans[(blk + 1) * NR + lane];
            cn[1] = (lane + 32 < NR) ? trans[(blk + 1) * NR + lane + 32] : 0.0f;
        }
        const float *f = fwd + (blk + rd) * NSTATE;            // forward vector before this block
        const float *b = bwd + (blk + rd + 1) * NSTATE;        // backward vector after it
        float x[2];
        x[0] = (c0v + b[to[0]]) + f[from[0]];
        x[1] = (lane + 32 < NR) ? (c1v + b[to[1]]) + f[from[1]] : -INFINITY;
        float m = fmaxf(x[0], x[1]);
        m = seg_max<32>(m);
        float sum = expf(x[0] - m) + ((x[1] > -INFINITY) ? expf(x[1] - m) : 0.0f);
        sum = seg_sum<32>(sum);
        const float lse = m + logf(sum);
        float *pc = tpost + blk * NR;
        pc[lane] = x[0] - lse;
        if (lane + 32 < NR) pc[lane + 32] = x[1] - lse;
    }
}

// trace (decode.c:499-543) from LOG posteriors: one thread per (read-local) trace row.
// OutT = uint8_t: the batched path (what fast5_interface.c:126-143 writes), saturating at 255; OutT = int32_t: the
// trace_from_posterior drop-in, which keeps the reference's int (a posterior sum a shade above 1 rounds to 256 there).
template <int NBASE, bool IS_LOG, typename OutT>
__global__ void trace_kernel(const float *__restrict__ tpost, const int64_t *__restrict__ blk_off, int n_reads,
                             OutT *__restrict__ trace) {
    constexpr int NSTATE = 2 * NBASE;
    constexpr int NR = NSTATE * (NBASE + 1);
    const int rd = blockIdx.y;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    if (T <= 0) return;
    const float *tp = tpost + b0 * NR;
    OutT *tr = trace + (b0 + rd) * NSTATE;
    auto ex = [](float v) { return IS_LOG ? expf(v) : v; };
    auto quant = [](float sum) -> OutT {
        const int v = (int)roundf(255.0f * sum);
        if (sizeof(OutT) == 1) return (OutT)min(255, max(0, v));
        return (OutT)v;
    };
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row <= T; row += gridDim.x * blockDim.x) {
        OutT *o = tr + (int64_t)row * NSTATE;
        if (row == 0) {
            // mass LEAVING each state in block 0 (decode.c:511-518)
            for (int from = 0; from < NSTATE; from++) {
                float sum = 0.0f;
                for (int to = 0; to < NBASE; to++) sum += ex(tp[to * NSTATE + from]);
                sum += ex(tp[NBASE * NSTATE + from]);
                o[from] = quant(sum);
            }
        } else {
            const float *pc = tp + (int64_t)(row - 1) * NR;
            for (int to = 0; to < NBASE; to++) {
                float sum = ex(pc[to * NSTATE]);
                for (int from = 1; from < NSTATE; from++) sum += ex(pc[to * NSTATE + from]);
                o[to] = quant(sum);
            }
            const float *pf = pc + NBASE * NSTATE;
            for (int to = NBASE; to < NSTATE; to++) {
                const float sum = ex(pf[to - NBASE]) + ex(pf[to]);
                o[to] = quant(sum);
            }
        }
    }
}

__global__ void exp_inplace_kernel(float *__restrict__ x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = expf(x[i]);
}

}  // namespace ffb

static inline int nbase_of(int nr) { return nr == 40 ? 4 : (nr == 60 ? 5 : 0); }
#define FFB_OKL(n) (cudaGetLastError() == cudaSuccess ? (n) : -1)

int ffb_launch_logz(const float *trans, const int64_t *blk_off, int n_reads, int nr, double *logZ, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    switch (nbase_of(nr)) {
    case 4: ffb::logz_kernel<4><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, logZ); break;
    case 5: ffb::logz_kernel<5><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, logZ); break;
    default: return -1;
    }
    return FFB_OKL(1);
}

int ffb_launch_sub_logz(float *trans, const int64_t *blk_off, int n_reads, int nr, const double *logZ,
                        int64_t total_blocks, cudaStream_t st) {
    if (n_reads <= 0 || total_blocks <= 0) return 0;
    // x: chunks within a read, y: reads (y <= 65535 per launch)
    int launches = 0;
    for (int r0 = 0; r0 < n_reads; r0 += 65535) {
        const int nr_reads = (n_reads - r0) < 65535 ? (n_reads - r0) : 65535;
        dim3 grid(8, nr_reads);
        ffb::sub_logz_kernel<<<grid, 256, 0, st>>>(trans, blk_off + r0, nr_reads, nr, logZ + r0);
        launches++;
    }
    return FFB_OKL(launches);
}

int ffb_launch_viterbi(const float *trans, const int64_t *blk_off, int n_reads, int nr, uint64_t *tb_scratch,
                       int32_t *path, float *qpath, float *score, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    switch (nbase_of(nr)) {
    case 4: ffb::viterbi_kernel<4><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, tb_scratch, path, qpath, score); break;
    case 5: ffb::viterbi_kernel<5><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, tb_scratch, path, qpath, score); break;
    default: return -1;
    }
    return FFB_OKL(1);
}

// fwd_scratch: 2 * (total_blocks + n_reads) * nstate floats (forward and backward state vectors)
int ffb_launch_transpost(const float *trans, const int64_t *blk_off, int n_reads, int nr, float *fwd_scratch,
                         float *tpost, int64_t total_blocks, cudaStream_t st) {
    if (n_reads <= 0 || total_blocks <= 0) return 0;
    const int nbase = nbase_of(nr);
    if (nbase != 4 && nbase != 5) return -1;
    float *bwd_scratch = fwd_scratch + (total_blocks + n_reads) * 2 * nbase;
    if (getenv("FFB_FB_SEQUENTIAL")) {      // the two-pass version (forward scan, then backward scan + normalisation)
        if (nbase == 4) {
            ffb::fb_fwd_kernel<4><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, fwd_scratch);
            ffb::fb_bwd_kernel<4><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, fwd_scratch, tpost);
        } else {
            ffb::fb_fwd_kernel<5><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, fwd_scratch);
            ffb::fb_bwd_kernel<5><<<n_reads, 32, 0, st>>>(trans, blk_off, n_reads, fwd_scratch, tpost);
        }
        return FFB_OKL(2);
    }
    const int64_t warps = (total_blocks < 148 * 64) ? total_blocks : 148 * 64;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    if (nbase == 4) {
        ffb::fb_scan2_kernel<4><<<n_reads, 64, 0, st>>>(trans, blk_off, n_reads, fwd_scratch, bwd_scratch);
        ffb::fb_combine_kernel<4><<<grid, 256, 0, st>>>(trans, blk_off, n_reads, fwd_scratch, bwd_scratch, tpost, total_blocks);
    } else {
        ffb::fb_scan2_kernel<5><<<n_reads, 64, 0, st>>>(trans, blk_off, n_reads, fwd_scratch, bwd_scratch);
        ffb::fb_combine_kernel<5><<<grid, 256, 0, st>>>(trans, blk_off, n_reads, fwd_scratch, bwd_scratch, tpost, total_blocks);
    }
    return FFB_OKL(2);
}

template <typename OutT>
static int launch_trace_t(const float *tpost, const int64_t *blk_off, int n_reads, int nr, OutT *trace, int is_log, cudaStream_t st) {
    int launches = 0;
    for (int r0 = 0; r0 < n_reads; r0 += 65535) {
        const int nrd = (n_reads - r0) < 65535 ? (n_reads - r0) : 65535;
        dim3 grid(8, nrd);
        // blk_off is absolute, so shifting the pointer keeps per-read addressing intact except
        // for the "+ rd" row padding of the trace, handled by passing the shifted trace base.
        const int nb = nbase_of(nr);
        OutT *tb = trace + (int64_t)r0 * 2 * nb;
        if (nb == 4 && is_log) ffb::trace_kernel<4, true, OutT><<<grid, 128, 0, st>>>(tpost, blk_off + r0, nrd, tb);
        else if (nb == 4) ffb::trace_kernel<4, false, OutT><<<grid, 128, 0, st>>>(tpost, blk_off + r0, nrd, tb);
        else if (nb == 5 && is_log) ffb::trace_kernel<5, true, OutT><<<grid, 128, 0, st>>>(tpost, blk_off + r0, nrd, tb);
        else if (nb == 5) ffb::trace_kernel<5, false, OutT><<<grid, 128, 0, st>>>(tpost, blk_off + r0, nrd, tb);
        else return -1;
        launches++;
    }
    return FFB_OKL(launches);
}

// wide = 0: one saturating byte per entry (batched path); wide = 1: int32 per entry, the reference's own values
int ffb_launch_trace(const float *tpost, const int64_t *blk_off, int n_reads, int nr, void *trace, int is_log, int wide,
                     cudaStream_t st) {
    if (n_reads <= 0) return 0;
    return wide ? launch_trace_t(tpost, blk_off, n_reads, nr, (int32_t *)trace, is_log, st)
                : launch_trace_t(tpost, blk_off, n_reads, nr, (uint8_t *)trace, is_log, st);
}

int ffb_launch_exp_inplace(float *x, int64_t n, cudaStream_t st) {
    if (n <= 0) return 0;
    ffb::exp_inplace_kernel<<<148 * 8, 256, 0, st>>>(x, n);
    return FFB_OKL(1);
}
