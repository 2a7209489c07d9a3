// rnn_tc.cu -- the recurrent hot loop on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces reference grumod_forward/backward + grumod_step (src/layers.c:571-715) and
// lstm_forward/backward + lstm_step (src/layers.c:877-1026).
//
// Decomposition.  A thread-block cluster of C CTAs has up to GMAX SLOTS; each slot works through a
// list of GROUPS of 16 whole reads, one group after the other (the host deals the length-sorted
// groups longest-first to the least-loaded slot, so a ragged or oversized batch keeps every slot
// busy for about the same number of steps and the whole layer is ONE wave of co-resident clusters);
// CTA c owns hidden units [c*HS, (c+1)*HS), HS = S/C = 32.
// Per group and step the CTA computes
//       a[gate g, hidden j][read] = sum_k sW[g*S + j][k] * h_{t-1}[read][k]
// as one M=128 x N=16 x K=S tensor-core GEMM:
//   A = the CTA's slice of sW, RESIDENT IN TENSOR MEMORY for all T steps (fp16 hi/lo planes,
//       2 x 128 of the 512 TMEM columns; row = lane).  With A in shared memory every MMA re-reads
//       128 rows x 32 B = 4 KB of it (32 clk at 128 B/clk) whatever N is, which made the N=16
//       MMAs 4x slower than the tensor pipe; from TMEM they run at the pipe rate (8 clk).
//       Row order: TMEM lane quadrant q holds, for the 8 hidden units 8q..8q+7,
//       lanes [0,8) = gate 0, [8,16) = gate 1, [16,24) = gate 2, [24,32) = gate 3 (GRU: zero),
//       so a 16-lane x 256-bit tcgen05.ld hands every thread two gates of the SAME cell and the
//       gate arithmetic needs no shuffles.
//   B = the previous state of the group's 16 reads, [k-group][plane hi/lo][read][8 halfs],
//   D = fp32 in tensor memory.  fp32-faithful product: hi*hi + hi*lo + lo*hi (tc_common.cuh).
//       The tensor core truncates on every accumulate (tools/probe_acc.py), so the hi*hi
//       product is split over two K-halves into separate accumulators and the cross terms
//       into a third; the three are added in registers with round-to-nearest.
//
// The layer is a chain of T dependent steps, so what bounds it is the LATENCY of one step's
// chain (MMA -> gates -> state exchange across the cluster), not any throughput: the groups
// are software-pipelined against each other -- each has its own barriers, its own 4 gate
// warps (one per TMEM quadrant) and its own control warp -- so while one group's state is in
// flight another group's MMAs and gates run.
//
// State exchange.  Every CTA needs all S elements of h_t of a group.  Pushing slices through
// distributed shared memory measured 8-9 B/clk per CTA and cluster-scope fences ~1.3 us
// (profiles/r01_exchange_microbench_*.txt); instead each CTA bulk-stores its staged slice to
// a small L2-resident ring in global memory and issues ONE multicast bulk load that lands it
// in the B operand of all C CTAs and signals their mbarriers (~80 B/clk, no fences).
//
// Step protocol of one group (all waits are CTA-scope mbarrier waits):
//   control thread : wait h_full (all slices of h_{t-1} landed) -> arm h_full for h_t ->
//                    3*S/16 tcgen05.mma -> commit -> acc_full; when the MMAs have retired, tell
//                    every peer "I have consumed h_{t-1}" (remote arrive on its h_empty)
//   4 gate warps   : prefetch Xin_t, wait acc_full, tcgen05.ld, gates, blend with the register-
//                    resident previous state, write h_t to HBM (fp32 and/or fp16 hi/lo planes
//                    for the next layer's tensor GEMM), stage the slice -> arrive `staged`
//   control thread : wait staged -> bulk store slice to the ring -> wait h_empty (all peers
//                    consumed h_{t-1}) -> multicast load ring -> B of every CTA (-> their h_full)
// Gate order and arithmetic as the reference: grumod_step src/layers.c:664-715 (z, r, n);
// lstm_step src/layers.c:979-1026 (i, f, g, o).
#include <algorithm>

#include "ffb_common.cuh"
#include "tc_common.cuh"

namespace ffb {
using namespace tc;

// LO_SMEM: the lo plane of the weight slice lives in SHARED memory instead of tensor memory (S = 512: the two planes
// of a 128 x 512 slice would fill all 512 TMEM columns).  Its MMAs (Wlo * hhi) then fetch A from shared memory
// (~32 clk each instead of 9), which S = 512 pays for a third of its MMAs.
// NACC: accumulators per group.  3 (default): Whi*hhi split over the two K-halves + one for the cross terms.  2: all of
// Whi*hhi in one accumulator (S/16 truncating full-magnitude accumulations instead of S/32) + the cross terms -- used
// where the third accumulator would cost a whole slot of the cluster (S = 384: 4 slots instead of 2).
template <int S_, int C_, int NGATE_, bool LO_SMEM_ = false, int NACC_ = 3, int GCAP_ = 5>
struct RnnTcCfg {
    static constexpr int S = S_, C = C_, NGATE = NGATE_, NACC = NACC_;
    static constexpr bool LO_SMEM = LO_SMEM_;
    static constexpr int HS = S / C;                  // hidden units per CTA
    static constexpr int NQ = HS / 8;                 // TMEM quadrants in use = k-groups per slice
    static constexpr int KG = S / 8;                  // 16-byte k-groups along K
    static constexpr int NG = 16;                     // reads per group (MMA N)
    static constexpr int TMEM_COLS = 512;
    static constexpr int A_COLS = S / 2;              // TMEM columns of one A plane (two halfs per 32-bit column)
    static constexpr int A_PLANE = 128 * S * 2;       // bytes of one plane of the global weight image [128 rows][S halfs]
    static constexpr int A_SMEM = LO_SMEM ? A_PLANE : 0;   // shared-memory copy of the lo plane, [k-group][128 rows][16 B]
    static constexpr int ACC_COL0_ = LO_SMEM ? A_COLS : 2 * A_COLS;
    // groups per cluster: what TMEM holds next to the weight plane(s) -- and, with the lo plane in shared memory, what
    // fits next to it there
    // 5 slots x 5 warps = 800 threads leave 72 registers per thread; 6 slots (960 threads, 64 registers, spills in the
    // gate warps) measured slower: 34.9 vs 34.2 ms per step at S=256 although 16 more SMs went to the streamed GEMM
    static constexpr int GCAP = GCAP_;
    static constexpr int GMAX_T = (TMEM_COLS - ACC_COL0_) / (NACC * 16) < GCAP ? (TMEM_COLS - ACC_COL0_) / (NACC * 16) : GCAP;
    static constexpr int GMAX = LO_SMEM ? (GMAX_T < 2 ? GMAX_T : 2) : GMAX_T;
    static constexpr int LBO_B = 2 * NG * 16;         // bytes between k-groups of B (hi and lo planes interleaved)
    static constexpr int B_GROUP = KG * LBO_B;        // bytes of one group's B operand
    static constexpr int SLICE = NQ * LBO_B;          // bytes of one CTA's slice of one group's state
    static constexpr int ACC_COLS = NACC * NG;        // TMEM columns per group
    static constexpr int ACC_COL0 = ACC_COL0_;        // accumulators follow the A plane(s)
    static constexpr int WARPS_PER_GROUP = 5;         // 4 gate warps + 1 control warp
    static constexpr int MAX_THREADS = GMAX * WARPS_PER_GROUP * 32;
    static_assert(HS == 32, "four quadrants of 8 hidden units");
    static_assert(GMAX >= 1 && ACC_COL0 + GMAX * ACC_COLS <= TMEM_COLS, "tensor memory budget");
    static_assert(C <= 16, "cluster size");
    __host__ __device__ static constexpr size_t smem_bytes(int G) {
        return (size_t)A_SMEM + (size_t)G * (B_GROUP + SLICE) + 64;
    }
    __host__ __device__ static constexpr size_t ring_bytes(int n_clusters, int G) {
        return (size_t)n_clusters * 2 * G * C * SLICE;
    }
};

__device__ __forceinline__ void bulk_store_global(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// one read of global memory, delivered to the same shared-memory offset of every CTA in `mask`,
// completing `bytes` on the mbarrier at the same offset in each of them
__device__ __forceinline__ void bulk_load_multicast(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// CTA-scope remote arrive: orders nothing but the arrival itself (used for "operand consumed")
__device__ __forceinline__ void mbar_arrive_remote_cta(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// 16 lanes x 256 bit, two column blocks: thread t gets rows t/4 (v0,v1,v4,v5) and t/4 + 8
// (v2,v3,v6,v7), columns 2*(t%4)+{0,1} (v0..v3) and 8 + 2*(t%4)+{0,1} (v4..v7)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// Optional phase timing of one control thread and one gate warp (cluster 0, CTA 0, group 0):
// build with FFB_EXTRA_NVCC_FLAGS=-DFFB_RNN_PROFILE, read with ffb_test_rnn_prof() (testhooks.cu).
#ifdef FFB_RNN_PROFILE
// per-launch timeline of the profile build: globaltimer at entry / exit of CTA 0 of launch k at [2k] / [2k + 1] (tools/step_timeline.py)
__device__ unsigned long long ffb_rnn_tl[64];
__device__ unsigned ffb_rnn_tl_n;
__device__ __forceinline__ unsigned long long ffb_gtime() { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); return t_; }
__device__ unsigned long long ffb_rnn_prof_dev[16];
#define PROF_DECL unsigned long long pt_ = clock64(), pa_[12] = {0}; const bool prof_ = (blockIdx.x == 0 && g == 0)
#define PROF(i) do { const unsigned long long n_ = clock64(); pa_[i] += n_ - pt_; pt_ = n_; } while (0)
#define PROF_FLUSH(lo, hi) do { if (prof_) for (int i_ = lo; i_ < hi; i_++) ffb_rnn_prof_dev[i_] += pa_[i_]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_FLUSH(lo, hi)
#endif

// gate arithmetic of the cells: 0 = MUFU ex2 / rcp as they come, 1 = expf + IEEE division, 2 = MUFU with repaired argument
// scaling and one Newton step (ffb_common.cuh)
#ifndef FFB_RNN_GATES
#define FFB_RNN_GATES 0
#endif
// Timing-only ablations (results are garbage): which part of the step bounds the recurrence?  Bit mask, see
// tools/ablate_timing.py: 1 = cells without MUFU, 2 = a quarter of the MMAs, 4 = no Xin loads / no output stores,
// 8 = no cross-proxy fence after staging, 16 = no bulk store to the ring, 32 = no tcgen05.ld; the parts of 4 one by one:
// 64 = no Xin loads, 128 = no copy-out of the fp16 planes, 256 = no stores of the fused rows; 512 / 1024 = the fused-row
// stores / the Xin loads go to a window of 1024 rows (L2-resident: same instructions and requests, no HBM traffic);
// 2048 / 4096 = the Xin loads / the fused-row stores as 16-byte accesses (a quarter of the instructions, same bytes)
#ifndef FFB_RNN_ABLATE
#define FFB_RNN_ABLATE 0
#endif
#define XR(r) ((FFB_RNN_ABLATE & 1024) ? ((r) & 1023) : (r))
#define ZR(r) ((FFB_RNN_ABLATE & 512) ? ((r) & 1023) : (r))
__device__ __forceinline__ float gate_logistic(float x) {
#if FFB_RNN_ABLATE & 1
    return fminf(fmaxf(fmaf(x, 0.25f, 0.5f), 0.0f), 1.0f);
#elif FFB_RNN_GATES == 1
    return logisticf(x);
#elif FFB_RNN_GATES == 2
    return mid_logistic(x);
#else
    return fast_logistic(x);
#endif
}
__device__ __forceinline__ float gate_tanh(float x) {
#if FFB_RNN_ABLATE & 1
    return fminf(fmaxf(x, -1.0f), 1.0f);
#elif FFB_RNN_GATES == 1
    return tanh_ref(x);
#elif FFB_RNN_GATES == 2
    return mid_tanh(x);
#else
    return fast_tanh(x);
#endif
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::MAX_THREADS, 1)
rnn_tc_kernel(const float *__restrict__ Xin, const __half *__restrict__ Wimg, float *__restrict__ Hout,
              __half *__restrict__ Hhi, __half *__restrict__ Hlo, const int32_t *__restrict__ order,
              const int64_t *__restrict__ blk_off, const int32_t *__restrict__ slot_off, const int32_t *__restrict__ slot_list,
              uint8_t *__restrict__ ring, int *__restrict__ progress, int G, int n_groups, int backward,
              const float *__restrict__ bnext, float *__restrict__ xnext, int next_ld, int next_rows, float ff_scale) {
    constexpr int S = Cfg::S, C = Cfg::C, NGATE = Cfg::NGATE, NG = Cfg::NG, GMAX = Cfg::GMAX;
    // GRU only: the fourth gate slot of the weight image (rows 24..31 of every TMEM quadrant, zero otherwise) carries the
    // z-gate rows of the NEXT layer's input projection.  They ride in the same M=128 MMAs for free: at step s the slot
    // holds iW_z * h_{s-1}, i.e. the next layer's Xin[.][0..S) of the previous time index (see the header comment).
    // The TOP layer carries the flip-flop output layer there instead (next_rows = 40 / 60 rows of FF_W, the rest zero):
    // xnext = trans [blocks][next_ld], written as tanh(. + b) / ff_scale (globalnorm_manystay, src/layers.c:1082-1087).
    const bool fuse = (NGATE == 3) && xnext != nullptr;
    const bool fuse_ff = next_rows > 0;
    extern __shared__ __align__(128) uint8_t smem[];
#ifdef FFB_RNN_PROFILE
    __shared__ unsigned tl_k;
    if (blockIdx.x == 0 && threadIdx.x == 0) { tl_k = atomicAdd(&ffb_rnn_tl_n, 1u); if (tl_k < 32) ffb_rnn_tl[2 * tl_k] = ffb_gtime(); }
#endif
    __shared__ uint64_t h_full[GMAX], h_empty[GMAX], acc_full[GMAX], staged[GMAX];
    __shared__ uint32_t tmem_slot;

    uint8_t *A_lo_smem = smem;                                   // LO_SMEM only: [KG][128 rows][16 B]
    uint8_t *B_base = smem + Cfg::A_SMEM;                        // [G][KG][hi|lo][NG][16 B]
    uint8_t *stg_base = B_base + (size_t)G * Cfg::B_GROUP;       // [G][NQ][hi|lo][NG][16 B]

    const uint32_t crank = cluster_ctarank();
    const int cluster_id = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);    // provably warp-uniform: roles, TMEM and staging addresses stay in uniform registers
    const int nthreads = blockDim.x;

    // ---- one-time setup ----
    {
        uint4 *bz = reinterpret_cast<uint4 *>(B_base);            // h_{-1} = 0
        for (int i = tid; i < (int)(((size_t)G * Cfg::B_GROUP) / 16); i += nthreads) bz[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int g = 0; g < G; g++) {
            mbar_init(&h_full[g], 1);
            mbar_init(&h_empty[g], C);
            mbar_init(&acc_full[g], 1);
            mbar_init(&staged[g], Cfg::NQ);
        }
        fence_barrier_init();
    }
    if constexpr (Cfg::LO_SMEM) {
        // lo plane -> shared memory in the no-swizzle K-major operand layout: k-group kg, row r at (kg * 128 + r) * 16 B
        const uint4 *lo = reinterpret_cast<const uint4 *>(Wimg) + ((size_t)crank * 2 * Cfg::A_PLANE + Cfg::A_PLANE) / 16;
        for (int i = tid; i < 128 * Cfg::KG; i += nthreads) {
            const int kg = i % Cfg::KG, r = i / Cfg::KG;         // consecutive threads read consecutive 16-byte chunks of a row
            reinterpret_cast<uint4 *>(A_lo_smem)[kg * 128 + r] = lo[(size_t)r * Cfg::KG + kg];
        }
    }
    if (warp == 0) tmem_alloc(&tmem_slot, Cfg::TMEM_COLS);
    fence_proxy_async_smem();       // zeroed B (and the lo plane) were written through the generic proxy
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp < 4) {
        // weight slice -> tensor memory: this thread owns TMEM lane 32*warp + lane = row of the image
        const uint4 *src = reinterpret_cast<const uint4 *>(Wimg) + ((size_t)crank * 2 * Cfg::A_PLANE + (size_t)(warp * 32 + lane) * S * 2) / 16;
#pragma unroll 1
        for (int plane = 0; plane < (Cfg::LO_SMEM ? 1 : 2); plane++) {
            const uint4 *row = src + (size_t)plane * (Cfg::A_PLANE / 16);
            const uint32_t tdst = tmem + ((uint32_t)(warp * 32) << 16) + plane * Cfg::A_COLS;
#pragma unroll 4
            for (int c = 0; c < Cfg::A_COLS; c += 8) {
                const uint4 v0 = row[c / 4], v1 = row[c / 4 + 1];
                const uint32_t v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                tmem_st8(tdst + c, v);
            }
        }
        tmem_st_wait();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    cluster_sync_all();             // every CTA's barriers are initialised before any remote arrive / copy
    // this CTA is resident: a dependent grid (the next layer's streamed input GEMM) may be scheduled on the SMs
    // this grid leaves free -- it synchronises on `progress`, not on this grid's completion
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef FFB_RNN_PROFILE
    if (progress && blockIdx.x == 0 && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ffb_rnn_prof_dev[12] = t_; }
#endif

    const int g = (warp < 4 * G) ? (warp >> 2) : (warp - 4 * G);    // this warp's slot of the cluster
    // the slot walks a list of 16-read groups, one after the other (host: longest-first to the least-loaded slot)
    const int sl0 = slot_off[cluster_id * G + g], sl1 = slot_off[cluster_id * G + g + 1];
    uint8_t *Bg = B_base + (size_t)g * Cfg::B_GROUP;
    uint8_t *stg = stg_base + (size_t)g * Cfg::SLICE;
    const uint32_t acc = tmem + Cfg::ACC_COL0 + (uint32_t)g * Cfg::ACC_COLS;
    uint32_t gs = 0;                // steps this slot has run so far: barrier phases continue across groups

    if (warp >= 4 * G) {
        // =========================== control warp of group g ===========================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_f16(128, NG);
            const uint32_t tA_hi = tmem, tA_lo = tmem + Cfg::A_COLS;
            const uint64_t dB_hi = make_smem_desc(smem_u32(Bg), Cfg::LBO_B, 128, LAYOUT_NONE);
            const uint64_t dB_lo = make_smem_desc(smem_u32(Bg) + NG * 16, Cfg::LBO_B, 128, LAYOUT_NONE);
            const uint64_t dA_lo = make_smem_desc(smem_u32(A_lo_smem), 128 * 16, 128, LAYOUT_NONE);   // LO_SMEM only
            uint8_t *ring_g = ring + ((((size_t)cluster_id * 2) * G + g) * C + crank) * Cfg::SLICE;   // parity 0
            const size_t ring_par = (size_t)G * C * Cfg::SLICE;
            PROF_DECL;
            uint32_t ga = 0;        // acc_full completions so far (one per step, plus one per group when the z rows are fused)
            auto issue_mmas = [&]() {
                tcgen05_fence_after();
                // 3*S/16 MMAs into THREE accumulators -- for accuracy, not speed: N=16 MMAs with uniform-register
                // operands issue at ~9 clk whichever accumulator they target (profiles/r01_mma_indep_microbench.txt).
                // Within an accumulator the 2^-11-sized cross terms go first, so each one sees only S/32
                // full-magnitude (truncating) accumulations.
                //   chain 0: Whi*hlo then Whi*hhi over K-half 0     chain 1: the same over K-half 1
                //   chain 2: Wlo*hhi over all of K
                constexpr int KS = (FFB_RNN_ABLATE & 2) ? S / 64 : S / 16, KH = KS / 2;
                if constexpr (Cfg::NACC == 3) {
#pragma unroll
                for (int i = 0; i < KS; i++) {
                    const int k0 = i < KH ? i : i - KH, k1 = k0 + KH;            // k-steps of chains 0 / 1
                    const uint64_t ob0 = (uint64_t)((k0 * 2 * Cfg::LBO_B) >> 4), ob1 = (uint64_t)((k1 * 2 * Cfg::LBO_B) >> 4),
                                   ob2 = (uint64_t)((i * 2 * Cfg::LBO_B) >> 4);
                    const uint64_t b0 = (i < KH ? dB_lo : dB_hi) + ob0, b1 = (i < KH ? dB_lo : dB_hi) + ob1;
                    umma_f16_ts(acc, tA_hi + k0 * 8, b0, idesc, i != 0);
                    umma_f16_ts(acc + NG, tA_hi + k1 * 8, b1, idesc, i != 0);
                    if constexpr (Cfg::LO_SMEM)
                        umma_f16(acc + 2 * NG, dA_lo + (uint64_t)((i * 2 * 128 * 16) >> 4), dB_hi + ob2, idesc, i != 0);
                    else
                        umma_f16_ts(acc + 2 * NG, tA_lo + i * 8, dB_hi + ob2, idesc, i != 0);
                }
                } else if constexpr (Cfg::NACC == 2) {
                    // two accumulators: acc 0 = Whi*hhi over all of K, acc 1 = Whi*hlo + Wlo*hhi
#pragma unroll
                for (int i = 0; i < KS; i++) {
                    const uint64_t ob = (uint64_t)((i * 2 * Cfg::LBO_B) >> 4);
                    umma_f16_ts(acc + NG, tA_hi + i * 8, dB_lo + ob, idesc, i != 0);
                    umma_f16_ts(acc + NG, tA_lo + i * 8, dB_hi + ob, idesc, 1);
                    umma_f16_ts(acc, tA_hi + i * 8, dB_hi + ob, idesc, i != 0);
                }
                } else {
                    // ONE accumulator (as the input GEMM, gemm_tc.cu): the 2^-11-sized cross terms go in first -- their
                    // truncation is invisible -- and the S/16 full-magnitude Whi*hhi products on top: the same number of
                    // full-magnitude truncating accumulations as a dedicated accumulator, half the tcgen05.ld traffic
#pragma unroll
                for (int i = 0; i < KS; i++) {
                    const uint64_t ob = (uint64_t)((i * 2 * Cfg::LBO_B) >> 4);
                    umma_f16_ts(acc, tA_hi + i * 8, dB_lo + ob, idesc, i != 0);
                    umma_f16_ts(acc, tA_lo + i * 8, dB_hi + ob, idesc, 1);
                }
#pragma unroll
                for (int i = 0; i < KS; i++)
                    umma_f16_ts(acc, tA_hi + i * 8, dB_hi + (uint64_t)((i * 2 * Cfg::LBO_B) >> 4), idesc, 1);
                }
                umma_commit(&acc_full[g]);
            };
            for (int sl = sl0; sl < sl1; sl++) {
            // slots are sorted by length (descending): the group's first read is its longest
            const int rd0 = order[slot_list[sl] * NG];
            const int Tmax = rd0 >= 0 ? (int)(blk_off[rd0 + 1] - blk_off[rd0]) : 0;
            for (int s = 0; s < Tmax; s++, gs++, ga += (NGATE == 3)) {
                const uint32_t ph = gs & 1u;
                if (s > 0) mbar_wait(&h_full[g], ph ^ 1u);           // h_{s-1} complete in B
                PROF(0);
                mbar_arrive_expect_tx(&h_full[g], C * Cfg::SLICE);    // arm phase s: peers push h_s only after my MMA(s)
                if (s == 0) {
                    // h_{-1} = 0: nothing to multiply (and B still holds the previous group's last state) -- the gate
                    // warps take a = 0 for this step
                    mbar_arrive(&acc_full[g]);
                } else {
                    issue_mmas();
                }
                PROF(1);
                mbar_wait(&acc_full[g], (NGATE == 3 ? ga : gs) & 1u);
                PROF(2);
                // the MMAs have retired: this CTA no longer reads h_{s-1}
                for (uint32_t d = 0; d < (uint32_t)C; d++) mbar_arrive_remote_cta(&h_empty[g], d);
                mbar_wait(&staged[g], (NGATE == 3 ? ga : gs) & 1u);                       // the gate warps staged my slice of h_s
                PROF(3);
                uint8_t *rg = ring_g + (size_t)ph * ring_par;
#ifdef FFB_RNN_FENCE1
                // the gate warps' generic-proxy writes of the slice happen-before this point (syncwarp -> arrive.release ->
                // wait.acquire); ONE cross-proxy fence by the thread that issues the copy orders them before it
                fence_proxy_async_smem();
#endif
                if (!(FFB_RNN_ABLATE & 16)) bulk_store_global(rg, stg, Cfg::SLICE);
                PROF(4);
                mbar_wait(&h_empty[g], ph);                           // every peer has consumed h_{s-1}
                PROF(5);
                // no proxy fence: the ring was written by the async proxy (bulk store, completed by
                // wait_group) and is read by the async proxy
                bulk_load_multicast(Bg + crank * Cfg::SLICE, rg, Cfg::SLICE, &h_full[g], (uint16_t)((1u << C) - 1u));
                PROF(6);
            }
            // drain: the group's last copies still target this CTA; they must have landed before h_full is re-armed for
            // the next group and before anybody exits
            if (Tmax > 0) mbar_wait(&h_full[g], (gs - 1u) & 1u);
            if (fuse && Tmax > 0) {
                // one more round of MMAs on the group's LAST state: the next layer's z rows of the last time index.  It
                // has retired (acc_full) before this thread tells any peer that the operand may be overwritten.
                // `staged`: all four gate warps have read it out of tensor memory -- only then may acc_full move on (a warp
                // still waiting for this phase would miss it once the next group's first step completes the following one).
                issue_mmas();
                mbar_wait(&acc_full[g], (NGATE == 3 ? ga : gs) & 1u);
                mbar_wait(&staged[g], (NGATE == 3 ? ga : gs) & 1u);
                ga++;
            }
            }
            PROF_FLUSH(0, 7);
        }
        __syncwarp();
    } else {
        // =========================== gate warp (group g, quadrant q) ===========================
        const int q = warp & 3;
        const int e = lane >> 2, cp = lane & 3;
        const int j = crank * Cfg::HS + q * 8 + e;          // global hidden index of this thread's cells
        constexpr int XROW = NGATE * S;
        // this thread's four cells: reads col(i) = (i>>1)*8 + 2*cp + (i&1) of the group.  Per cell a
        // running pointer into Xin and a running output row, stepped by +-1 block per step.
        const int rstep = backward ? -1 : 1;
        const float *const xj = Xin + j;                    // this thread's column of the input projection; row = orow[i]
        // second role of the lane: after the slice is staged, copy one 16-byte chunk (8 hidden units of one
        // read, one plane) from the staging tile to the fp16 layer output -- off the critical path
        const int cl_read = lane & 15, cl_plane = lane >> 4;
        __half *cl_dst = (cl_plane ? Hlo : Hhi);
        if (cl_dst) cl_dst += crank * Cfg::HS + q * 8;
        const uint4 *cl_src = reinterpret_cast<const uint4 *>(stg + (size_t)q * Cfg::LBO_B + cl_plane * NG * 16 + cl_read * 16);

        const uint32_t t_lo = acc + ((uint32_t)(q * 32) << 16);       // lanes 32q .. 32q+15: gates 0, 1
        const uint32_t t_hi = acc + ((uint32_t)(q * 32 + 16) << 16);  // lanes 32q+16 .. 32q+31: gates 2, 3
        // staging: [kg_local = q][plane][read][8 halfs], this thread writes element e of its 4 reads
        __half *st_hi = reinterpret_cast<__half *>(stg + (size_t)q * Cfg::LBO_B) + e;
        __half *st_lo = st_hi + NG * 8;

#ifdef FFB_RNN_PROFILE
        unsigned long long pt_ = clock64(), pa_[12] = {0}; const bool prof_ = (blockIdx.x == 0 && warp == 0 && lane == 0);
#endif
        uint32_t ga = 0;            // acc_full phases seen so far (steps + fused drain rounds)
        const float bz = fuse ? bnext[j] : 0.0f;
        float *const xz = fuse ? xnext + j : nullptr;       // next layer's Xin: z-gate column of hidden unit j
        for (int sl = sl0; sl < sl1; sl++) {
        const int grp = slot_list[sl];                      // this round's group of 16 reads
        int cT[4];
        int32_t orow[4];            // the host guarantees total blocks < 2^31
        float hprev[4], cstate[4];
        int Tmin = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int col = (i >> 1) * 8 + 2 * cp + (i & 1);
            const int rd = order[grp * NG + col];
            cT[i] = rd >= 0 ? (int)(blk_off[rd + 1] - blk_off[rd]) : 0;
            const int32_t base = rd >= 0 ? (int32_t)blk_off[rd] : 0;
            orow[i] = base + ((backward && cT[i] > 0) ? cT[i] - 1 : 0);
            hprev[i] = 0.0f; cstate[i] = 0.0f;
            Tmin = min(Tmin, cT[i]);
        }
        Tmin = min(Tmin, __shfl_xor_sync(0xffffffffu, Tmin, 1));   // shortest read of the group: while s < Tmin
        Tmin = min(Tmin, __shfl_xor_sync(0xffffffffu, Tmin, 2));   // no cell needs a predicate
        int Tmax = 0;               // the group's first read is its longest (sorted)
        {
            const int rd0 = order[grp * NG];
            Tmax = rd0 >= 0 ? (int)(blk_off[rd0 + 1] - blk_off[rd0]) : 0;
        }
        int cl_T = 0;
        int32_t cl_row = 0;
        {
            const int rd = order[grp * NG + cl_read];
            cl_T = rd >= 0 ? (int)(blk_off[rd + 1] - blk_off[rd]) : 0;
            cl_row = (rd >= 0 ? (int32_t)blk_off[rd] : 0) + ((backward && cl_T > 0) ? cl_T - 1 : 0);
        }
        // the input projection of step sn, rows orow[i] + dr: issued a whole step ahead (right after this warp has staged
        // step sn - 1), because with a lead of only the other phases of the step (~1200 clk) the loads of the slowest of the
        // cluster's 32 gate warps were still in flight when its accumulators arrived -- ~300 clk per step and warp of
        // long-scoreboard stall at the first use of x, on the step's critical path (profiles/r02_rnn_tc_ncu_summary.txt)
        float x[4][NGATE];
        auto fetch_x = [&](int sn, int dr) {
            if (FFB_RNN_ABLATE & (4 | 64)) {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int gt = 0; gt < NGATE; gt++) x[i][gt] = 0.01f * (float)(i + gt);
            } else if (FFB_RNN_ABLATE & 2048) {
                // the same bytes per thread in NGATE 16-byte loads instead of 4 * NGATE 4-byte loads (wrong elements: timing only)
#pragma unroll
                for (int gt = 0; gt < NGATE; gt++) {
                    const float4 v = __ldcs(reinterpret_cast<const float4 *>(Xin + (int64_t)(orow[gt & 3] + dr) * XROW + gt * S + (j & ~3)));
                    x[0][gt] = v.x; x[1][gt] = v.y; x[2][gt] = v.z; x[3][gt] = v.w;
                }
            } else if (sn < Tmin) {                    // warp-uniform
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int gt = 0; gt < NGATE; gt++) x[i][gt] = __ldcs(xj + (int64_t)XR(orow[i] + dr) * XROW + gt * S);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int gt = 0; gt < NGATE; gt++) x[i][gt] = (sn < cT[i]) ? __ldcs(xj + (int64_t)XR(orow[i] + dr) * XROW + gt * S) : 0.0f;
            }
        };
#ifndef FFB_RNN_LATE_PREFETCH
        fetch_x(0, 0);
#endif
        for (int s = 0; s < Tmax; s++, gs++, ga += (NGATE == 3)) {
            const bool all = s < Tmin;          // warp-uniform
#ifdef FFB_RNN_LATE_PREFETCH
            fetch_x(s, 0);                      // A/B: the round-1 placement, at the top of the step
#endif
            PROF(8);
            mbar_wait(&acc_full[g], (NGATE == 3 ? ga : gs) & 1u);
            PROF(9);
            tcgen05_fence_after();
            // ---- TMEM -> registers: a[i][gate], three partial accumulators added round-to-nearest ----
            float a[4][4];
            if (s == 0 || (FFB_RNN_ABLATE & 32)) {
                // h_{-1} = 0 (layers.c:586 / :892 zero the initial state): no MMAs were issued for this step
#pragma unroll
                for (int i = 0; i < 4; i++) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.0f;
            } else if constexpr (Cfg::NACC == 3) {
                float v0[8], v1[8];
                tmem_ld_16x256b_x2(t_lo, v0);
                tmem_ld_16x256b_x2(t_lo + NG, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int r0 = (i >> 1) * 4 + (i & 1);
                    a[i][0] = v0[r0] + v1[r0]; a[i][1] = v0[r0 + 2] + v1[r0 + 2];
                }
                tmem_ld_16x256b_x2(t_hi, v0);
                tmem_ld_16x256b_x2(t_hi + NG, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int r0 = (i >> 1) * 4 + (i & 1);
                    a[i][2] = v0[r0] + v1[r0]; a[i][3] = v0[r0 + 2] + v1[r0 + 2];
                }
                tmem_ld_16x256b_x2(t_lo + 2 * NG, v0);
                tmem_ld_16x256b_x2(t_hi + 2 * NG, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int r0 = (i >> 1) * 4 + (i & 1);
                    a[i][0] = acc_comp(a[i][0]) + v0[r0]; a[i][1] = acc_comp(a[i][1]) + v0[r0 + 2];
                    a[i][2] = acc_comp(a[i][2]) + v1[r0]; a[i][3] = acc_comp(a[i][3]) + v1[r0 + 2];
                }
            } else if constexpr (Cfg::NACC == 1) {
                float v0[8], v2[8];
                tmem_ld_16x256b_x2(t_lo, v0);
                tmem_ld_16x256b_x2(t_hi, v2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int r0 = (i >> 1) * 4 + (i & 1);
                    a[i][0] = acc_comp(v0[r0]); a[i][1] = acc_comp(v0[r0 + 2]);
                    a[i][2] = acc_comp(v2[r0]); a[i][3] = acc_comp(v2[r0 + 2]);
                }
            } else {
                float v0[8], v1[8], v2[8], v3[8];
                tmem_ld_16x256b_x2(t_lo, v0);
                tmem_ld_16x256b_x2(t_lo + NG, v1);
                tmem_ld_16x256b_x2(t_hi, v2);
                tmem_ld_16x256b_x2(t_hi + NG, v3);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int r0 = (i >> 1) * 4 + (i & 1);
                    // accumulator 0 (Whi*hhi) took the full-magnitude truncating accumulations, accumulator 1 the small cross terms
                    a[i][0] = acc_comp(v0[r0]) + v1[r0]; a[i][1] = acc_comp(v0[r0 + 2]) + v1[r0 + 2];
                    a[i][2] = acc_comp(v2[r0]) + v3[r0]; a[i][3] = acc_comp(v2[r0 + 2]) + v3[r0 + 2];
                }
            }
            tcgen05_fence_before();
            PROF(10);
            // ---- cells ----
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float hn, cn = 0.0f;
                if constexpr (NGATE == 3) {
                    const float z = gate_logistic(x[i][0] + a[i][0]);                   // layers.c:697-699
                    const float r = gate_logistic(x[i][1] + a[i][1]);
                    const float hbar = gate_tanh(r * a[i][2] + x[i][2]);               // layers.c:704-709
                    hn = z * hprev[i] + (1.0f - z) * hbar;                             // layers.c:712-714
                } else {
                    const float ig = gate_logistic(x[i][0] + a[i][0]);                  // layers.c:1013-1024
                    const float fg = gate_logistic(x[i][1] + a[i][1]);
                    const float gg = gate_tanh(x[i][2] + a[i][2]);
                    const float og = gate_logistic(x[i][3] + a[i][3]);
                    cn = fg * cstate[i] + ig * gg;
                    hn = og * gate_tanh(cn);
                }
                if (all || s < cT[i]) {             // finished reads keep (and keep pushing) their frozen state
                    hprev[i] = hn;
                    if constexpr (NGATE == 4) cstate[i] = cn;
                }
                __half shv, slv;
                split_f16(hprev[i], shv, slv);
                const int col = (i >> 1) * 8 + 2 * cp + (i & 1);
                st_hi[col * 8] = shv;
                st_lo[col * 8] = slv;
            }
#if !defined(FFB_RNN_FENCE1) && !(FFB_RNN_ABLATE & 8)
            fence_proxy_async_smem();      // staged slice -> visible to the bulk-copy engine
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(&staged[g]);
            PROF(11);
#ifndef FFB_RNN_LATE_PREFETCH
            // ---- the NEXT step's input projection (x is dead from here on) ----
            if (s + 1 < Tmax) fetch_x(s + 1, rstep);
#endif
            // ---- layer output (not on the step's critical path) ----
            if (!(FFB_RNN_ABLATE & (4 | 128)) && cl_dst && s < cl_T) *reinterpret_cast<uint4 *>(cl_dst + (int64_t)cl_row * S) = *cl_src;
            __syncwarp();       // the staging tile has been read (by other lanes than those that rewrite it next step)
            if (!(FFB_RNN_ABLATE & 4) && Hout) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (all || s < cT[i]) __stcs(Hout + (int64_t)orow[i] * S + j, hprev[i]);
            }
            if (!(FFB_RNN_ABLATE & (4 | 256)) && fuse && s > 0) {
                // gate slot 3 = iW_z(next layer) * h_{s-1}: the next layer's z pre-activation of the PREVIOUS time index
                if (!fuse_ff && (FFB_RNN_ABLATE & 4096)) {
                    // one 16-byte store per thread instead of four 4-byte stores (wrong places: timing only)
                    if (all || s <= cT[0])
                        __stcs(reinterpret_cast<float4 *>(xnext + (j & ~3) + (int64_t)(orow[0] - rstep) * next_ld), make_float4(a[0][3] + bz, a[1][3], a[2][3], a[3][3]));
                } else if (!fuse_ff) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (all || s <= cT[i]) __stcs(xz + (int64_t)ZR(orow[i] - rstep) * next_ld, a[i][3] + bz);
                } else if (j < next_rows) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (all || s <= cT[i]) xz[(int64_t)(orow[i] - rstep) * next_ld] = tanh_ref(a[i][3] + bz) / ff_scale;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) orow[i] += rstep;
            cl_row += rstep;
            // ---- publish progress for the consumer of the output planes ----
            if (progress && (((s + 1) % FFB_RNN_PUBLISH_PERIOD) == 0 || s == Tmax - 1)) {
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(progress + grp, 1);
            }
        }
        if (fuse && Tmax > 0) {
            // fused drain round: the z rows of the group's last time index (only reads that ran all Tmax steps still owe it)
            mbar_wait(&acc_full[g], (NGATE == 3 ? ga : gs) & 1u);
            ga++;
            tcgen05_fence_after();
            float v0[8], v1[8], a3[4];
            tmem_ld_16x256b_x2(t_hi, v0);
            if constexpr (Cfg::NACC > 1) tmem_ld_16x256b_x2(t_hi + NG, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; i++) a3[i] = acc_comp(v0[(i >> 1) * 4 + (i & 1) + 2]) + (Cfg::NACC > 1 ? v1[(i >> 1) * 4 + (i & 1) + 2] : 0.0f);
            if constexpr (Cfg::NACC == 3) {
                tmem_ld_16x256b_x2(t_hi + 2 * NG, v0);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; i++) a3[i] += v0[(i >> 1) * 4 + (i & 1) + 2];
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&staged[g]);
            if (!fuse_ff) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cT[i] == Tmax) __stcs(xz + (int64_t)(orow[i] - rstep) * next_ld, a3[i] + bz);
            } else if (j < next_rows) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cT[i] == Tmax) xz[(int64_t)(orow[i] - rstep) * next_ld] = tanh_ref(a3[i] + bz) / ff_scale;
            }
        }
        }
        PROF_FLUSH(8, 12);
    }

    tcgen05_fence_before();
    __syncthreads();
    // every gate warp fenced its last stores before its final publish: count this CTA as finished (the
    // coarse dependency of GEMM tiles that span more than four reads)
    if (progress && tid == 0) atomicAdd(progress + n_groups, 1);
#ifdef FFB_RNN_PROFILE
    if (progress && blockIdx.x == 0 && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ffb_rnn_prof_dev[13] = t_; }
#endif
    cluster_sync_all();
    if (warp == 0) tmem_dealloc(tmem, Cfg::TMEM_COLS);
#ifdef FFB_RNN_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0 && tl_k < 32) ffb_rnn_tl[2 * tl_k + 1] = ffb_gtime();
#endif
}

// two accumulators everywhere TMEM holds both weight planes: four tcgen05.ld instead of six on the step's critical path
// (S=256: -4.5 % per layer, trans deviation 4.8e-5 vs 4.6e-5 with three accumulators, tools/report_parity.py)
#ifndef FFB_RNN_NACC
#define FFB_RNN_NACC 2
#endif
using GruTc256 = RnnTcCfg<256, 8, 3, false, FFB_RNN_NACC>;
using LstmTc256 = RnnTcCfg<256, 8, 4, false, FFB_RNN_NACC>;
// six slots per cluster (960 threads, 64 registers): more groups in flight per SM for batches that do not fit one wave
using GruTc256x6 = RnnTcCfg<256, 8, 3, false, FFB_RNN_NACC, 6>;
using LstmTc256x6 = RnnTcCfg<256, 8, 4, false, FFB_RNN_NACC, 6>;
// S = 384: 12-CTA clusters (non-portable size), 32 hidden units per CTA again; 2 x 192 TMEM columns of weights leave 128 for
// the accumulators.  ONE accumulator per group (cross terms first, then the full-magnitude Whi*hhi products: the same number
// of full-magnitude truncating accumulations as a dedicated accumulator, profiles/r02_ab_notes.txt) makes room for FIVE slots
// instead of four: 7 clusters x 5 = 35 slots take the 64 groups of a 1024-read batch in two rounds instead of three.
#ifndef FFB_RNN_NACC384
#define FFB_RNN_NACC384 1
#endif
using GruTc384 = RnnTcCfg<384, 12, 3, false, FFB_RNN_NACC384>;
using LstmTc384 = RnnTcCfg<384, 12, 4, false, FFB_RNN_NACC384>;
using GruTc512 = RnnTcCfg<512, 16, 3, true>;    // r103_native: 16-CTA clusters, hi plane in tensor memory (256 columns),
using LstmTc512 = RnnTcCfg<512, 16, 4, true>;   // lo plane in shared memory (128 KB), 2 groups

}  // namespace ffb

// ---------------------------------------------------------------------------------------
// shape dispatch: f is a generic lambda called with a value-initialised config object
// wide = true: the instance with the most slots per cluster (S = 256: six, 960 threads at 64 registers); launches with
// G <= 5 use the five-slot instance (72 registers)
template <class F>
static auto tc_dispatch(int kind, int S, F &&f, bool wide = false) {
    if (S == 384) return kind == 0 ? f(ffb::GruTc384{}) : f(ffb::LstmTc384{});
    if (S == 512) return kind == 0 ? f(ffb::GruTc512{}) : f(ffb::LstmTc512{});
    if (wide) return kind == 0 ? f(ffb::GruTc256x6{}) : f(ffb::LstmTc256x6{});
    return kind == 0 ? f(ffb::GruTc256{}) : f(ffb::LstmTc256{});
}

int ffb_rnn_tc_prof(unsigned long long *out, int reset) {
#ifdef FFB_RNN_PROFILE
    if (out && cudaMemcpyFromSymbol(out, ffb::ffb_rnn_prof_dev, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ffb::ffb_rnn_prof_dev, z, sizeof z); }
    return 1;
#else
    (void)out; (void)reset;
    return 0;
#endif
}
int ffb_rnn_tc_timeline(unsigned long long *out, int reset) {      // out[64]; returns the number of launches stamped, -1 without the profile build
#ifdef FFB_RNN_PROFILE
    unsigned n = 0;
    if (cudaMemcpyFromSymbol(&n, ffb::ffb_rnn_tl_n, sizeof n) != cudaSuccess) return -1;
    if (out && cudaMemcpyFromSymbol(out, ffb::ffb_rnn_tl, 64 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) { unsigned z = 0; unsigned long long zz[64] = {0}; cudaMemcpyToSymbol(ffb::ffb_rnn_tl_n, &z, sizeof z); cudaMemcpyToSymbol(ffb::ffb_rnn_tl, zz, sizeof zz); }
    return (int)n;
#else
    (void)out; (void)reset;
    return -1;
#endif
}
int ffb_rnn_tc_supported(int kind, int S) { return (kind == 0 || kind == 1) && (S == 256 || S == 384 || S == 512); }
int ffb_rnn_tc_cluster_size(int kind, int S) { return tc_dispatch(kind, S, [](auto cfg) { return (int)decltype(cfg)::C; }); }
int ffb_rnn_tc_rmax(int kind, int S) { return tc_dispatch(kind, S, [](auto cfg) { return (int)(decltype(cfg)::GMAX * decltype(cfg)::NG); }, true); }

size_t ffb_rnn_tc_image_halfs(int kind, int S) {
    return tc_dispatch(kind, S, [](auto cfg) { using Cfg = decltype(cfg); return (size_t)Cfg::C * 2 * Cfg::A_PLANE / 2; });
}

size_t ffb_rnn_tc_ring_bytes(int kind, int S, int n_clusters, int R) {
    return tc_dispatch(kind, S, [&](auto cfg) { return decltype(cfg)::ring_bytes(n_clusters, R / 16); }, true);
}

// sW [G*S][S] (row per output) -> per-CTA tensor-memory images (fp16 bit patterns):
// [cta][plane hi/lo][row = 32*quad + 8*gate + e][S halfs]; GRU rows with gate 3 are zero, or -- fused_z [S][S] given --
// row jj of the next layer's input projection (its z gate), computed for free by the same MMAs
template <class Cfg>
static void pack_image(const float *sW, uint16_t *img, const float *fused_z) {
    constexpr int S = Cfg::S;
    const size_t plane_halfs = Cfg::A_PLANE / 2;
    for (size_t i = 0; i < (size_t)Cfg::C * 2 * plane_halfs; i++) img[i] = 0;
    for (int c = 0; c < Cfg::C; c++) {
        uint16_t *hi = img + (size_t)c * 2 * plane_halfs, *lo = hi + plane_halfs;
        for (int q = 0; q < Cfg::NQ; q++)
            for (int g = 0; g < (Cfg::NGATE == 3 && fused_z ? 4 : Cfg::NGATE); g++)
                for (int e = 0; e < 8; e++) {
                    const int jj = c * Cfg::HS + q * 8 + e, row = 32 * q + 8 * g + e;
                    for (int k = 0; k < S; k++) {
                        const float w = g < Cfg::NGATE ? sW[(size_t)(g * S + jj) * S + k] : fused_z[(size_t)jj * S + k];
                        const __half h = __float2half_rn(w);
                        const __half l = __float2half_rn(w - __half2float(h));
                        hi[(size_t)row * S + k] = __half_as_ushort(h);
                        lo[(size_t)row * S + k] = __half_as_ushort(l);
                    }
                }
    }
}
void ffb_rnn_tc_pack(int kind, int S, const float *sW, uint16_t *img, const float *fused_z) {
    tc_dispatch(kind, S, [&](auto cfg) { pack_image<decltype(cfg)>(sW, img, fused_z); return 0; });
}
// GRU: a quarter of every CTA's M=128 rows is free -- room for S rows of the next layer's input projection
int ffb_rnn_tc_can_fuse_z(int kind, int S) { return kind == 0 && ffb_rnn_tc_supported(kind, S); }

template <class Cfg>
static int prepare_one() {
    if (cudaFuncSetAttribute(ffb::rnn_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes(Cfg::GMAX)) != cudaSuccess) return -1;
    if (Cfg::C > 8 && cudaFuncSetAttribute(ffb::rnn_tc_kernel<Cfg>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -1;
    return 0;
}
int ffb_rnn_tc_prepare(int kind, int S) {
    if (!ffb_rnn_tc_supported(kind, S)) return -1;
    if (tc_dispatch(kind, S, [](auto cfg) { return prepare_one<decltype(cfg)>(); }, true) != 0) return -1;
    return tc_dispatch(kind, S, [](auto cfg) { return prepare_one<decltype(cfg)>(); });
}

template <class Cfg>
static void rnn_tc_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int n_clusters, int G, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(n_clusters * Cfg::C);
    cfg.blockDim = dim3(G * Cfg::WARPS_PER_GROUP * 32);
    cfg.dynamicSmemBytes = Cfg::smem_bytes(G);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
}

// how many clusters of this kernel can be co-resident (0 on error)
template <class Cfg>
static int max_clusters_one(int G) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    rnn_tc_config<Cfg>(cfg, attr, 64, G, 0);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ffb::rnn_tc_kernel<Cfg>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int ffb_rnn_tc_max_clusters(int kind, int S, int R) {
    if (!ffb_rnn_tc_supported(kind, S)) return 0;
    return tc_dispatch(kind, S, [&](auto cfg) { return max_clusters_one<decltype(cfg)>(std::min(R / 16, (int)decltype(cfg)::GMAX)); }, true);
}

template <class Cfg>
static int launch_one(const float *Xin, const void *Wimg, float *Hout, void *Hhi, void *Hlo, const RnnBatch &rb,
                      const RnnTcSched &sched, int R, int backward, void *ring, int *progress, const float *bnext, float *xnext,
                      int next_rows, float ff_scale, cudaStream_t st) {
    const int G = R / Cfg::NG;
    if (G < 1 || G > Cfg::GMAX || R % Cfg::NG || !ring || !sched.slot_off || !sched.slot_list) return -1;
    const int n_clusters = sched.n_clusters;
    if (n_clusters == 0) return 0;
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    rnn_tc_config<Cfg>(cfg, attr, n_clusters, G, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, ffb::rnn_tc_kernel<Cfg>, Xin, (const __half *)Wimg, Hout, (__half *)Hhi,
                                       (__half *)Hlo, rb.order, rb.blk_off, sched.slot_off, sched.slot_list, (uint8_t *)ring, progress,
                                       G, sched.n_groups, backward, bnext, xnext, next_rows > 0 ? next_rows : Cfg::NGATE * Cfg::S,
                                       next_rows, ff_scale);
    return e == cudaSuccess ? 1 : -1;
}

int ffb_launch_rnn_tc(int kind, int S, const float *Xin, const void *Wimg, float *Hout, void *Hhi, void *Hlo,
                      const RnnBatch &rb, const RnnTcSched &sched, int R, int backward, void *ring, int *progress,
                      const float *bnext, float *xnext, int next_rows, float ff_scale, cudaStream_t st) {
    if (!ffb_rnn_tc_supported(kind, S)) return -1;
    if (xnext && (!bnext || !ffb_rnn_tc_can_fuse_z(kind, S) || next_rows < 0 || next_rows > S)) return -1;
    return tc_dispatch(kind, S, [&](auto cfg) { return launch_one<decltype(cfg)>(Xin, Wimg, Hout, Hhi, Hlo, rb, sched, R, backward, ring, progress, bnext, xnext, next_rows, ff_scale, st); },
                       R / 16 > 5);
}
