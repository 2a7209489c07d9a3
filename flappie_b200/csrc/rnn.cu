// rnn.cu -- the recurrent hot loop: one grumod / LSTM layer over a ragged batch of whole reads.
//
// Replaces reference grumod_forward/backward + grumod_step (src/layers.c:571-715) and
// lstm_forward/backward + lstm_step (src/layers.c:877-1026), which run one cblas_sgemv +
// SSE gate arithmetic per time step per read (75-82 % of the reference's run time).
//
// B200 mapping (fp32 CUDA-core path):
//   * A thread-block CLUSTER of C CTAs owns R reads for the whole layer.  CTA `c` of the
//     cluster owns hidden units [c*S/C, (c+1)*S/C): its slice of the recurrent weights
//     sW (S x G*S/C fp32, 96-147 KB) is loaded ONCE and stays resident in shared memory for
//     all T steps, next to the full previous state H [S][R].
//   * per step: every CTA computes its [R x G*S/C] slice of  a = H_{t-1} * sW  from
//     shared memory (4 reads x 2 hidden x G gates of accumulators per thread), adds the
//     precomputed input projection Xin_t (prefetched from HBM under the k-loop), applies
//     the gates, writes h_t to HBM (the layer output) and broadcasts its h_t slice into
//     the H buffers of all C CTAs through distributed shared memory.  Two cluster barriers
//     per step order "everyone finished reading H_{t-1}" / "every h_t slice has landed".
//   * reads are sorted by length; a read shorter than the cluster's longest simply stops
//     updating (its state is never read again).  Backward layers start every read at its
//     own last block, so forward and backward both run s = 0 .. T_r-1 with t = s or
//     T_r-1-s; padding never touches state.
//   * LSTM cell state lives in registers of the owning thread for the whole layer.
//
// Gate order: GRU (z, r, n) -- layers.c:697-714; LSTM (i, f, g, o) -- layers.c:1013-1024.
#include <cooperative_groups.h>

#include <type_traits>

#include "ffb_common.cuh"

namespace cg = cooperative_groups;

namespace ffb {

// KRES_: rows k of the weight slice held in shared memory; rows [KRES, S) are read from the packed image in global
// memory (L2-resident) every step -- only S = 512, whose 256 KB slice does not fit next to H.  Same k order, same bits.
template <int G_, int S_, int C_, int R_, int KRES_ = S_>
struct RnnCfg {
    static constexpr int G = G_, S = S_, C = C_, R = R_, KRES = KRES_;
    static constexpr int HS = S / C;          // hidden units per CTA
    static constexpr int NC = G * HS;         // weight columns per CTA
    static constexpr int NHP = HS / 2;        // hidden pairs
    static constexpr int WARPS = (R / 32) * (NHP / 4);
    static constexpr int THREADS = WARPS * 32;
    static constexpr size_t SMEM_W = (size_t)KRES * NC * sizeof(float);
    static constexpr size_t SMEM_H = (size_t)S * R * sizeof(float);
    static constexpr size_t SMEM = SMEM_W + SMEM_H;
    static_assert(S % C == 0 && HS % 8 == 0 && R % 32 == 0, "bad recurrent tiling");
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// packed weights: [C][S (k)][NHP][G][2]
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
rnn_layer_kernel(const float *__restrict__ Xin, const float *__restrict__ Wp, float *__restrict__ Hout,
                 const int32_t *__restrict__ order, const int64_t *__restrict__ blk_off, int backward) {
    constexpr int G = Cfg::G, S = Cfg::S, C = Cfg::C, R = Cfg::R, HS = Cfg::HS, NC = Cfg::NC;
    extern __shared__ __align__(16) float smem[];
    constexpr int KRES = Cfg::KRES;
    float *Ws = smem;                    // [KRES][NC]
    float *Hs = smem + (size_t)KRES * NC; // [S][R]

    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / C;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int rq = (lane & 7) + 8 * (warp % (R / 32));        // read quad
    const int hp = (lane >> 3) + 4 * (warp / (R / 32));       // hidden pair within the CTA slice

    // resident weights + zero initial state (layers.c:592 / :638 / :892 / :902)
    {
        const float4 *src = reinterpret_cast<const float4 *>(Wp + (size_t)crank * S * NC);
        float4 *dst = reinterpret_cast<float4 *>(Ws);
        for (int i = tid; i < KRES * NC / 4; i += Cfg::THREADS) dst[i] = src[i];
        float4 *hz = reinterpret_cast<float4 *>(Hs);
        for (int i = tid; i < S * R / 4; i += Cfg::THREADS) hz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // this thread's four reads
    int Tr[4];
    int64_t base[4];
    int Tmax_cluster = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int slot = cluster_id * R + rq * 4 + i;
        const int rd = order[slot];
        if (rd >= 0) {
            base[i] = blk_off[rd];
            Tr[i] = (int)(blk_off[rd + 1] - blk_off[rd]);
        } else {
            base[i] = 0;
            Tr[i] = 0;
        }
    }
    {   // slots are sorted by length (descending): slot 0 of the cluster is the longest read
        const int rd0 = order[cluster_id * R];
        Tmax_cluster = rd0 >= 0 ? (int)(blk_off[rd0 + 1] - blk_off[rd0]) : 0;
    }
    const int j0 = crank * HS + 2 * hp;   // first of this thread's two hidden units (global index)

    float cst[4][2];   // LSTM cell state (unused for GRU)
#pragma unroll
    for (int i = 0; i < 4; i++) cst[i][0] = cst[i][1] = 0.0f;

    // DSMEM views of every CTA's H buffer
    float *Hremote[C];
#pragma unroll
    for (int d = 0; d < C; d++) Hremote[d] = cluster.map_shared_rank(Hs, d);

    cluster.sync();   // weights + zeroed H visible cluster-wide before the first remote write

    for (int s = 0; s < Tmax_cluster; s++) {
        // ---- prefetch this step's input projection (consumed after the k-loop) ----
        float2 xv[4][G];
        bool act[4];
        int64_t row[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            act[i] = s < Tr[i];
            const int t = backward ? (Tr[i] - 1 - s) : s;
            row[i] = base[i] + t;
            if (act[i]) {
                const float *xp = Xin + row[i] * (int64_t)(G * S) + j0;
#pragma unroll
                for (int g = 0; g < G; g++) xv[i][g] = __ldcs(reinterpret_cast<const float2 *>(xp + g * S));
            } else {
#pragma unroll
                for (int g = 0; g < G; g++) xv[i][g] = make_float2(0.f, 0.f);
            }
        }

        // ---- a = H_{t-1} * sW slice ----
        float acc[4][2 * G];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int q = 0; q < 2 * G; q++) acc[i][q] = 0.0f;
        const float *hptr = Hs + 4 * rq;
        const float *wptr = Ws + hp * 2 * G;
        auto kstep = [&](int k, const float *wk, auto from_global) {
            const float4 hv = *reinterpret_cast<const float4 *>(hptr + k * R);
            float w[2 * G];
            if constexpr (G == 4) {
                float4 w0, w1;
                if constexpr (decltype(from_global)::value) {
                    w0 = __ldg(reinterpret_cast<const float4 *>(wk));
                    w1 = __ldg(reinterpret_cast<const float4 *>(wk + 4));
                } else {
                    w0 = *reinterpret_cast<const float4 *>(wk);
                    w1 = *reinterpret_cast<const float4 *>(wk + 4);
                }
                w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
                w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
            } else {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const float2 wg = *reinterpret_cast<const float2 *>(wk + 2 * g);
                    w[2 * g] = wg.x; w[2 * g + 1] = wg.y;
                }
            }
            const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int q = 0; q < 2 * G; q++) acc[i][q] = fmaf(h4[i], w[q], acc[i][q]);
        };
#pragma unroll 8
        for (int k = 0; k < KRES; k++) kstep(k, wptr + k * NC, std::false_type{});
        if constexpr (KRES < S) {
            static_assert(KRES == S || G == 4, "the L2-streamed tail exists for the LSTM only");
            const float *wglob = Wp + (size_t)crank * S * NC + hp * 2 * G;
#pragma unroll 8
            for (int k = KRES; k < S; k++) kstep(k, wglob + (size_t)k * NC, std::true_type{});
        }

        // previous state of this thread's own (read, hidden) cells, before anyone overwrites H
        float hprev[4][2];
        if constexpr (G == 3) {
            const float4 p0 = *reinterpret_cast<const float4 *>(Hs + (size_t)j0 * R + 4 * rq);
            const float4 p1 = *reinterpret_cast<const float4 *>(Hs + (size_t)(j0 + 1) * R + 4 * rq);
            hprev[0][0] = p0.x; hprev[1][0] = p0.y; hprev[2][0] = p0.z; hprev[3][0] = p0.w;
            hprev[0][1] = p1.x; hprev[1][1] = p1.y; hprev[2][1] = p1.z; hprev[3][1] = p1.w;
        }
        cluster_arrive();   // (1) this CTA has finished reading H_{t-1}

        // ---- gates ----
        float hn[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if constexpr (G == 3) {
                    // grumod_step, layers.c:697-714; acc index = 2*gate + e, gates (z, r, n)
                    const float xz = e ? xv[i][0].y : xv[i][0].x;
                    const float xr = e ? xv[i][1].y : xv[i][1].x;
                    const float xn = e ? xv[i][2].y : xv[i][2].x;
                    const float z = logisticf(xz + acc[i][0 + e]);
                    const float r = logisticf(xr + acc[i][2 + e]);
                    const float hbar = tanh_ref(r * acc[i][4 + e] + xn);
                    hn[i][e] = z * hprev[i][e] + (1.0f - z) * hbar;
                } else {
                    // lstm_step, layers.c:1013-1024; gates (i, f, g, o)
                    const float xi = e ? xv[i][0].y : xv[i][0].x;
                    const float xf = e ? xv[i][1].y : xv[i][1].x;
                    const float xg = e ? xv[i][2].y : xv[i][2].x;
                    const float xo = e ? xv[i][3].y : xv[i][3].x;
                    const float forget = logisticf(xf + acc[i][2 + e]) * cst[i][e];
                    const float update = logisticf(xi + acc[i][0 + e]) * tanh_ref(xg + acc[i][4 + e]);
                    const float cnew = forget + update;
                    if (act[i]) cst[i][e] = cnew;
                    hn[i][e] = logisticf(xo + acc[i][6 + e]) * tanh_ref(cnew);
                }
            }
            if (act[i]) {
                __stcs(reinterpret_cast<float2 *>(Hout + row[i] * (int64_t)S + j0), make_float2(hn[i][0], hn[i][1]));
            }
        }

        cluster_wait();     // (1) every CTA of the cluster has finished reading H_{t-1}

        // ---- broadcast this thread's h_t cells into all C copies of H ----
        const bool all_act = act[0] && act[1] && act[2] && act[3];
        if (all_act) {
            const float4 v0 = make_float4(hn[0][0], hn[1][0], hn[2][0], hn[3][0]);
            const float4 v1 = make_float4(hn[0][1], hn[1][1], hn[2][1], hn[3][1]);
#pragma unroll
            for (int d = 0; d < C; d++) {
                float *hd = Hremote[d] + (size_t)j0 * R + 4 * rq;
                *reinterpret_cast<float4 *>(hd) = v0;
                *reinterpret_cast<float4 *>(hd + R) = v1;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (!act[i]) continue;
#pragma unroll
                for (int d = 0; d < C; d++) {
                    float *hd = Hremote[d] + (size_t)j0 * R + 4 * rq + i;
                    hd[0] = hn[i][0];
                    hd[R] = hn[i][1];
                }
            }
        }
        cluster_arrive();   // (2) my slice of h_t is written everywhere
        cluster_wait();     // (2) all slices have landed
    }
    // no CTA may exit while peers can still write into its shared memory: the last
    // barrier pair above already guarantees that (every remote write precedes arrive (2)).
}

// ---- configurations --------------------------------------------------------------
using GruCfg256 = RnnCfg<3, 256, 8, 64>;    // r941_5mC, r10C_pcr, north-star r941_native: 96 KB W + 64 KB H
using GruCfg96 = RnnCfg<3, 96, 2, 64>;      // small shapes for tests
using GruCfg64 = RnnCfg<3, 64, 2, 64>;
using LstmCfg256 = RnnCfg<4, 256, 8, 64>;   // r941_rna002: 128 KB W + 64 KB H
using LstmCfg384 = RnnCfg<4, 384, 16, 32>;  // r941_native @4de542f: 144 KB W + 48 KB H, 16-CTA cluster
using LstmCfg512 = RnnCfg<4, 512, 16, 32, 256>;  // r103_native: cross-check path only -- half of the 256 KB slice in shared
                                                 // memory (128 KB W + 64 KB H), the other half re-read from L2 every step
using LstmCfg128 = RnnCfg<4, 128, 4, 64>;
using LstmCfg96 = RnnCfg<4, 96, 2, 32>;

template <class Cfg>
static int prepare_cfg() {
    auto kern = rnn_layer_kernel<Cfg>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM) != cudaSuccess) return -1;
    if (Cfg::C > 8) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -1;
    }
    return 0;
}

template <class Cfg>
static int launch_cfg(const float *Xin, const float *Wp, float *Hout, const RnnBatch &rb, int backward,
                      cudaStream_t st) {
    if (rb.n_slots % Cfg::R != 0) return -1;
    const int n_clusters = rb.n_slots / Cfg::R;
    if (n_clusters == 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * Cfg::C);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rnn_layer_kernel<Cfg>, Xin, Wp, Hout, rb.order, rb.blk_off, backward);
    return e == cudaSuccess ? 1 : -1;
}

template <class Cfg>
static void pack_cfg(const float *sW, float *packed) {
    // packed[c][k][hp][g][e] = sW[g*S + c*HS + 2*hp + e][k]
    for (int c = 0; c < Cfg::C; c++)
        for (int k = 0; k < Cfg::S; k++)
            for (int hp = 0; hp < Cfg::NHP; hp++)
                for (int g = 0; g < Cfg::G; g++)
                    for (int e = 0; e < 2; e++) {
                        const int j = c * Cfg::HS + 2 * hp + e;
                        packed[(((size_t)c * Cfg::S + k) * Cfg::NHP + hp) * (2 * Cfg::G) + 2 * g + e] =
                            sW[(size_t)(g * Cfg::S + j) * Cfg::S + k];
                    }
}

}  // namespace ffb

#define FFB_RNN_DISPATCH(kind, S, EXPR_GRU256, EXPR_GRU96, EXPR_GRU64, EXPR_L256, EXPR_L384, EXPR_L512, EXPR_L128, EXPR_L96, DEFAULT) \
    do {                                                                                                                \
        if ((kind) == 0 && (S) == 256) { EXPR_GRU256; }                                                                 \
        else if ((kind) == 0 && (S) == 96) { EXPR_GRU96; }                                                              \
        else if ((kind) == 0 && (S) == 64) { EXPR_GRU64; }                                                              \
        else if ((kind) == 1 && (S) == 256) { EXPR_L256; }                                                              \
        else if ((kind) == 1 && (S) == 384) { EXPR_L384; }                                                              \
        else if ((kind) == 1 && (S) == 512) { EXPR_L512; }                                                              \
        else if ((kind) == 1 && (S) == 128) { EXPR_L128; }                                                              \
        else if ((kind) == 1 && (S) == 96) { EXPR_L96; }                                                                \
        else { DEFAULT; }                                                                                               \
    } while (0)

int ffb_rnn_supported(int kind, int S) {
    FFB_RNN_DISPATCH(kind, S, return 1, return 1, return 1, return 1, return 1, return 1, return 1, return 1, return 0);
}

int ffb_rnn_reads_per_cluster(int kind, int S) {
    using namespace ffb;
    FFB_RNN_DISPATCH(kind, S, return GruCfg256::R, return GruCfg96::R, return GruCfg64::R, return LstmCfg256::R,
                     return LstmCfg384::R, return LstmCfg512::R, return LstmCfg128::R, return LstmCfg96::R, return 0);
}

size_t ffb_rnn_packed_floats(int kind, int S) { return (size_t)(kind == 0 ? 3 : 4) * S * S; }

void ffb_rnn_pack_weights(int kind, int S, const float *sW, float *packed) {
    using namespace ffb;
    FFB_RNN_DISPATCH(kind, S, pack_cfg<GruCfg256>(sW, packed), pack_cfg<GruCfg96>(sW, packed),
                     pack_cfg<GruCfg64>(sW, packed), pack_cfg<LstmCfg256>(sW, packed),
                     pack_cfg<LstmCfg384>(sW, packed), pack_cfg<LstmCfg512>(sW, packed), pack_cfg<LstmCfg128>(sW, packed),
                     pack_cfg<LstmCfg96>(sW, packed), (void)0);
}

int ffb_rnn_prepare(int kind, int S) {
    using namespace ffb;
    FFB_RNN_DISPATCH(kind, S, return prepare_cfg<GruCfg256>(), return prepare_cfg<GruCfg96>(),
                     return prepare_cfg<GruCfg64>(), return prepare_cfg<LstmCfg256>(),
                     return prepare_cfg<LstmCfg384>(), return prepare_cfg<LstmCfg512>(), return prepare_cfg<LstmCfg128>(),
                     return prepare_cfg<LstmCfg96>(), return -1);
}

int ffb_launch_rnn(int kind, int S, const float *Xin, const float *sW_packed, float *Hout, const RnnBatch &rb,
                   int backward, cudaStream_t st) {
    using namespace ffb;
    FFB_RNN_DISPATCH(kind, S, return launch_cfg<GruCfg256>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<GruCfg96>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<GruCfg64>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<LstmCfg256>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<LstmCfg384>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<LstmCfg512>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<LstmCfg128>(Xin, sW_packed, Hout, rb, backward, st),
                     return launch_cfg<LstmCfg96>(Xin, sW_packed, Hout, rb, backward, st), return -1);
}
