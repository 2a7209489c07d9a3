// rnn_tc.cu -- the recurrent hot loop on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same cluster decomposition as rnn.cu (a cluster of C CTAs owns R whole reads for the whole
// layer, CTA c owns hidden units [c*32, (c+1)*32)), but the per-step product
//       a[gate g, hidden j][read] = sum_k sW[g*S + j][k] * h_{t-1}[read][k]
// is one M=128 x N=R x K=S tensor-core GEMM per CTA per step:
//   A = this CTA's slice of sW, RESIDENT in shared memory for all T steps (fp16 hi/lo planes,
//       no-swizzle K-major).  Row order: TMEM lane quadrant q holds, for the 8 hidden units
//       8q..8q+7, lanes [0,8) = z gate, [8,16) = r gate, [16,24) = n gate, [24,32) unused --
//       so every warp (= quadrant = SM sub-partition) owns complete cells and the gate
//       arithmetic is spread evenly over all four sub-partitions with warp shuffles only.
//   B = the previous state of the cluster's R reads, fp16 hi/lo planes [read][k] K-major,
//   D = fp32 in tensor memory.  fp32-faithful product: hi*hi + hi*lo + lo*hi (tc_common.cuh).
//       The tensor core truncates (round-toward-zero) on every accumulate (measured:
//       ~ -0.6e-7 relative per MMA, tests/probe_acc.py), so the hi*hi product is split over
//       two K-halves into separate accumulators and the small cross terms into a third; the
//       three are added in registers with round-to-nearest.
// Step protocol (mbarrier based, no cluster-wide barrier on the critical path):
//   control thread : wait h_full (every slice of h_{t-1} has landed in B) -> issue 3*S/16
//                    tcgen05.mma -> tcgen05.commit -> acc_full; then tell every peer
//                    "I have consumed h_{t-1}" (remote arrive on its h_empty)
//   all 8 warps    : tcgen05.ld their quadrant/half, regroup (z, r, n) per cell with shuffles,
//                    add the prefetched input projection Xin_t, gates, blend with the
//                    register-resident previous state, write h_t to HBM (fp32 and/or fp16 hi/lo
//                    planes for the next layer's tensor GEMM), stage the slice in the B layout
//   control thread : wait h_empty (all peers consumed h_{t-1}) -> one cp.async.bulk per
//                    plane per peer pushes the slice into every CTA's B operand through
//                    distributed shared memory, completing bytes on the peer's h_full.
// Gate order and arithmetic as the reference: grumod_step src/layers.c:664-715.
#include "ffb_common.cuh"
#include "tc_common.cuh"

namespace ffb {
using namespace tc;

template <int S_, int C_>
struct GruTcCfg {
    static constexpr int S = S_, C = C_, G = 3;
    static constexpr int HS = S / C;                 // hidden units per CTA
    static constexpr int KG = S / 8;                 // 16-byte k-groups along K
    static constexpr int A_RG = 15;                  // 8-row groups stored per k-group (the 16th is never read back)
    static constexpr int LBO_A = A_RG * 128;         // bytes between k-groups of A
    static constexpr int A_PLANE = KG * LBO_A;       // bytes per A plane
    static constexpr int RMAX = 80;
    static constexpr int CELLS = RMAX / 8;           // cells per thread: (R/2 reads per half) / 4 lane groups
    static constexpr int NACC = 3;                   // hi*hi K-half 0, hi*hi K-half 1, cross terms
    static constexpr int THREADS = 256;
    static_assert(HS == 32, "four quadrants of 8 hidden units");
    __host__ __device__ static constexpr size_t smem_bytes(int R) {
        return 2 * (size_t)A_PLANE + 128              // A hi / lo (+ the aliased 16th row group of the last k-group)
               + 2 * (size_t)KG * R * 16              // B hi / lo
               + 4 * (size_t)(HS / 8) * R * 16        // staged slice hi / lo, double buffered
               + (size_t)R * 12 + 64;                 // per-read base row (int64) + length (int32)
    }
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gru_tc_kernel(const float *__restrict__ Xin, const __half *__restrict__ Wimg, float *__restrict__ Hout,
              __half *__restrict__ Hhi, __half *__restrict__ Hlo, const int32_t *__restrict__ order,
              const int64_t *__restrict__ blk_off, int R, int backward) {
    constexpr int S = Cfg::S, C = Cfg::C, HS = Cfg::HS, KG = Cfg::KG, CELLS = Cfg::CELLS;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t h_full, h_empty, acc_full;
    __shared__ uint32_t tmem_slot;

    uint8_t *A_hi = smem;
    uint8_t *A_lo = A_hi + Cfg::A_PLANE;
    uint8_t *B_hi = A_lo + Cfg::A_PLANE + 128;
    const uint32_t b_plane = (uint32_t)KG * R * 16;
    uint8_t *B_lo = B_hi + b_plane;
    uint8_t *stg_base = B_lo + b_plane;                               // 2 buffers x {hi, lo} x [HS/8][R][16 B]
    const uint32_t stg_plane = (uint32_t)(HS / 8) * R * 16;
    int64_t *rd_base = reinterpret_cast<int64_t *>(stg_base + 4 * stg_plane);   // [R]
    int32_t *rd_T = reinterpret_cast<int32_t *>(rd_base + R);                    // [R]

    const uint32_t crank = cluster_ctarank();
    const int cluster_id = blockIdx.x / C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3;          // TMEM lane quadrant: hidden units 8*quad .. 8*quad+7 of this CTA
    const int half = warp >> 2;         // which half of the reads this warp handles
    const int Rh = R / 2;
    const int e = lane & 7;             // hidden unit within the quadrant
    const int sub = lane >> 3;          // this lane's cells are reads c0 + 4*ci + sub

    // ---- one-time setup ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(Wimg) + (size_t)crank * (2 * Cfg::A_PLANE / 16);
        uint4 *dst = reinterpret_cast<uint4 *>(A_hi);
        for (int i = tid; i < 2 * Cfg::A_PLANE / 16; i += Cfg::THREADS) dst[i] = src[i];
        uint4 *bz = reinterpret_cast<uint4 *>(A_lo + Cfg::A_PLANE);   // the 128-byte tail + B: h_{-1} = 0
        for (int i = tid; i < (int)((128 + 2 * b_plane) / 16); i += Cfg::THREADS) bz[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < R; i += Cfg::THREADS) {
            const int rd = order[cluster_id * R + i];
            rd_base[i] = rd >= 0 ? blk_off[rd] : 0;
            rd_T[i] = rd >= 0 ? (int)(blk_off[rd + 1] - blk_off[rd]) : 0;
        }
    }
    if (tid == 0) {
        mbar_init(&h_full, 1);
        mbar_init(&h_empty, C);
        mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async_smem();       // A / zeroed B were written through the generic proxy
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    cluster_sync_all();             // every CTA's barriers are initialised before any remote arrive / copy
    const uint32_t tmem = tmem_slot;

    int Tmax = 0;
    {
        const int rd0 = order[cluster_id * R];
        Tmax = rd0 >= 0 ? (int)(blk_off[rd0 + 1] - blk_off[rd0]) : 0;
    }
    const int j = crank * HS + quad * 8 + e;         // global hidden index of this lane's cells
    const uint32_t idesc = make_idesc_f16(128, (uint32_t)R);
    const uint32_t slice_bytes = stg_plane;          // bytes pushed per plane per peer per step
    const uint32_t step_tx = 2u * slice_bytes * C;   // bytes landing in this CTA's B per step
    const int c0 = half * Rh;

    // per-cell constants: read slot, length, base row (cells beyond Rh/4 are inactive)
    float hprev[CELLS];
    int cT[CELLS];
    int64_t cbase[CELLS];
#pragma unroll
    for (int ci = 0; ci < CELLS; ci++) {
        hprev[ci] = 0.0f;
        const int rs = c0 + 4 * ci + sub;
        const bool ok = (4 * ci + sub) < Rh;
        cT[ci] = ok ? rd_T[rs] : 0;
        cbase[ci] = ok ? rd_base[rs] : 0;
    }

    for (int s = 0; s < Tmax; s++) {
        const uint32_t ph = (uint32_t)s & 1u;
        // staging is double buffered: the bulk copies of step s may still be reading their source
        // while step s+1 is staged; reuse at s+2 is safe because reaching it needs h_full(s+1)
        // here, which needs every peer's MMA(s+1), which needs my step-s push to have landed
        uint8_t *stg_hi = stg_base + (size_t)ph * 2 * stg_plane;
        uint8_t *stg_lo = stg_hi + stg_plane;

        // ---------------- control: MMA issue ----------------
        if (warp == 0) {
            if (elect_one()) {
                if (s > 0) mbar_wait_cluster(&h_full, ph ^ 1u);       // h_{s-1} complete in B (phase s-1)
                mbar_arrive_expect_tx(&h_full, step_tx);              // arm phase s: peers push h_s only after my MMA(s)
                tcgen05_fence_after();
                const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
                const uint32_t lbo_b = (uint32_t)R * 16;
#pragma unroll 4
                for (int ks = 0; ks < S / 16; ks++) {
                    const uint64_t dah = make_smem_desc(a_hi + ks * 2 * Cfg::LBO_A, Cfg::LBO_A, 128, LAYOUT_NONE);
                    const uint64_t dal = make_smem_desc(a_lo + ks * 2 * Cfg::LBO_A, Cfg::LBO_A, 128, LAYOUT_NONE);
                    const uint64_t dbh = make_smem_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128, LAYOUT_NONE);
                    const uint64_t dbl = make_smem_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128, LAYOUT_NONE);
                    const int kh = ks / (S / 32);                                       // K-half
                    umma_f16(tmem + kh * R, dah, dbh, idesc, (ks % (S / 32)) != 0);   // hi*hi
                    umma_f16(tmem + 2 * R, dah, dbl, idesc, ks != 0);                 // cross terms
                    umma_f16(tmem + 2 * R, dal, dbh, idesc, 1);
                }
                umma_commit(&acc_full);
            }
            __syncwarp();
        }

        // ---------------- all warps: prefetch this step's input projection ----------------
        float xz[CELLS], xr[CELLS], xn[CELLS];
#pragma unroll
        for (int ci = 0; ci < CELLS; ci++) {
            xz[ci] = xr[ci] = xn[ci] = 0.0f;
            if (s < cT[ci]) {
                const int t = backward ? (cT[ci] - 1 - s) : s;
                const float *xp = Xin + (cbase[ci] + t) * (int64_t)(3 * S) + j;
                xz[ci] = __ldcs(xp);
                xr[ci] = __ldcs(xp + S);
                xn[ci] = __ldcs(xp + 2 * S);
            }
        }

        mbar_wait(&acc_full, ph);
        tcgen05_fence_after();
        // once the MMAs have retired this CTA no longer reads h_{s-1}: tell every peer
        if (warp == 0 && lane < C) mbar_arrive_remote(&h_empty, (uint32_t)lane);

        // ---------------- TMEM -> registers, regroup (z, r, n) per cell ----------------
        float az[CELLS], ar[CELLS], an[CELLS];
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + c0;
#pragma unroll
        for (int cc = 0; cc < Cfg::RMAX / 2; cc += 8) {
            if (cc < Rh) {
                float a0[8], a1[8], a2[8];
                tmem_ld8(taddr + cc, a0);
                tmem_ld8(taddr + R + cc, a1);
                tmem_ld8(taddr + 2 * R + cc, a2);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float a = (a0[q] + a1[q]) + a2[q];      // round-to-nearest sum of the partial accumulators
                    const float tz = __shfl_sync(0xffffffffu, a, e);
                    const float tr = __shfl_sync(0xffffffffu, a, 8 + e);
                    const float tn = __shfl_sync(0xffffffffu, a, 16 + e);
                    if ((q & 3) == sub) {
                        az[(cc + q) >> 2] = tz; ar[(cc + q) >> 2] = tr; an[(cc + q) >> 2] = tn;
                    }
                }
            }
        }
        tcgen05_fence_before();

        // ---------------- cells ----------------
        __half *shi = reinterpret_cast<__half *>(stg_hi + ((size_t)quad * R + c0 + sub) * 16) + e;
        __half *slo = reinterpret_cast<__half *>(stg_lo + ((size_t)quad * R + c0 + sub) * 16) + e;
#pragma unroll
        for (int ci = 0; ci < CELLS; ci++) {
            if (4 * ci < Rh) {
                const float z = logisticf(xz[ci] + az[ci]);                         // layers.c:697-699
                const float r = logisticf(xr[ci] + ar[ci]);
                const float hbar = tanh_ref(r * an[ci] + xn[ci]);                  // layers.c:704-709
                const float hn = z * hprev[ci] + (1.0f - z) * hbar;               // layers.c:712-714
                if (s < cT[ci]) {
                    hprev[ci] = hn;
                    const int t = backward ? (cT[ci] - 1 - s) : s;
                    const int64_t row = cbase[ci] + t;
                    if (Hout) __stcs(Hout + row * S + j, hn);
                    if (Hhi) {
                        __half hi, lo;
                        split_f16(hn, hi, lo);
                        Hhi[row * S + j] = hi;
                        Hlo[row * S + j] = lo;
                    }
                }
                __half shv, slv;
                split_f16(hprev[ci], shv, slv);      // finished reads keep pushing their frozen state
                shi[(size_t)ci * 32] = shv;           // 4 reads = 4 * 16 bytes = 32 halfs apart
                slo[(size_t)ci * 32] = slv;
            }
        }
        fence_proxy_async_smem();      // staged slice -> visible to the bulk-copy engine
        __syncthreads();

        // ---------------- control: push the slice to every CTA of the cluster ----------------
        if (warp == 0) {
            if (elect_one()) {
                mbar_wait_cluster(&h_empty, ph);      // every peer has consumed h_{s-1}
                const uint32_t dst_off = crank * slice_bytes;
                for (uint32_t d = 0; d < (uint32_t)C; d++) {
                    dsmem_bulk_copy(B_hi + dst_off, stg_hi, slice_bytes, &h_full, d);
                    dsmem_bulk_copy(B_lo + dst_off, stg_lo, slice_bytes, &h_full, d);
                }
            }
            __syncwarp();
        }
    }

    // drain: the last step's copies still target peers; nobody may exit before they have landed
    if (Tmax > 0 && warp == 0 && elect_one()) mbar_wait_cluster(&h_full, (uint32_t)(Tmax - 1) & 1u);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

using GruTc256 = GruTcCfg<256, 8>;

}  // namespace ffb

// ---------------------------------------------------------------------------------------
int ffb_rnn_tc_supported(int kind, int S) { return kind == 0 && S == 256; }
int ffb_rnn_tc_rmax(int kind, int S) { (void)kind; (void)S; return ffb::GruTc256::RMAX; }

size_t ffb_rnn_tc_image_halfs(int kind, int S) {
    (void)kind; (void)S;
    return (size_t)ffb::GruTc256::C * 2 * ffb::GruTc256::A_PLANE / 2;
}

// sW [3S][S] (row per output) -> per-CTA shared-memory images (fp16 bit patterns):
// [cta][plane hi/lo][k-group][row group rg = 4*quad + gate (3 = zero pad)][row e][8 halfs]
void ffb_rnn_tc_pack(int kind, int S, const float *sW, uint16_t *img) {
    (void)kind;
    using Cfg = ffb::GruTc256;
    const size_t plane_halfs = Cfg::A_PLANE / 2;
    for (size_t i = 0; i < (size_t)Cfg::C * 2 * plane_halfs; i++) img[i] = 0;
    for (int c = 0; c < Cfg::C; c++) {
        uint16_t *hi = img + (size_t)c * 2 * plane_halfs, *lo = hi + plane_halfs;
        for (int kg = 0; kg < Cfg::KG; kg++)
            for (int q = 0; q < 4; q++)
                for (int g = 0; g < 3; g++)
                    for (int e = 0; e < 8; e++)
                        for (int x = 0; x < 8; x++) {
                            const int jj = c * Cfg::HS + q * 8 + e;
                            const float w = sW[(size_t)(g * S + jj) * S + kg * 8 + x];
                            const __half h = __float2half_rn(w);
                            const __half l = __float2half_rn(w - __half2float(h));
                            const int rg = 4 * q + g;
                            const size_t off = ((size_t)kg * Cfg::LBO_A + rg * 128 + e * 16) / 2 + x;
                            hi[off] = __half_as_ushort(h);
                            lo[off] = __half_as_ushort(l);
                        }
    }
}

int ffb_rnn_tc_prepare(int kind, int S) {
    if (!ffb_rnn_tc_supported(kind, S)) return -1;
    using Cfg = ffb::GruTc256;
    if (cudaFuncSetAttribute(ffb::gru_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes(Cfg::RMAX)) != cudaSuccess) return -1;
    return 0;
}

static void rnn_tc_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int n_clusters, int R, cudaStream_t st) {
    using Cfg = ffb::GruTc256;
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(n_clusters * Cfg::C);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::smem_bytes(R);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
}

// how many clusters of this kernel can be co-resident (0 on error)
int ffb_rnn_tc_max_clusters(int kind, int S, int R) {
    if (!ffb_rnn_tc_supported(kind, S)) return 0;
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    rnn_tc_config(cfg, attr, 64, R, 0);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ffb::gru_tc_kernel<ffb::GruTc256>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ffb_launch_rnn_tc(int kind, int S, const float *Xin, const void *Wimg, float *Hout, void *Hhi, void *Hlo,
                      const RnnBatch &rb, int R, int backward, cudaStream_t st) {
    if (!ffb_rnn_tc_supported(kind, S)) return -1;
    using Cfg = ffb::GruTc256;
    if (R < 16 || R > Cfg::RMAX || R % 16 || rb.n_slots % R) return -1;
    const int n_clusters = rb.n_slots / R;
    if (n_clusters == 0) return 0;
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    rnn_tc_config(cfg, attr, n_clusters, R, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, ffb::gru_tc_kernel<Cfg>, Xin, (const __half *)Wimg, Hout, (__half *)Hhi,
                                       (__half *)Hlo, rb.order, rb.blk_off, R, backward);
    return e == cudaSuccess ? 1 : -1;
}
