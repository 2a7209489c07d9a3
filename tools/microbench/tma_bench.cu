// tma_bench.cu -- how fast can one SM pull activation tiles into shared memory?
// (design input for the input-projection GEMM, flappie_b200/csrc/gemm_tc.cu; not part of the library)
//
// A [M][256] fp16 row-major (one plane).  Each CTA streams 64-row tiles through a ring of stages;
// a consumer thread just waits for each stage and frees it (no math).  Modes:
//   0: cp.async.bulk.tensor 2-D, box 64 cols x 64 rows, 128-byte swizzle (what gemm_ws does), 1 box per stage
//   1: same, but 4 boxes (the whole K=256 of the tile) per stage
//   2: plain cp.async.bulk of the tile's contiguous 32 KB (64 rows x 512 B), one copy per stage
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../flappie_b200/csrc/tc_common.cuh"

using namespace ffb::tc;

__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap map, const __half *A, int64_t n_tiles, int mode,
                                                     int stage_bytes, long long *cycles, int share, int stagger) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    const int loads_per_tile = mode == 0 ? 4 : 1;   // stages consumed per tile
    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t0i = blockIdx.x / share; t0i < n_tiles; t0i += gridDim.x / share) {
                const int64_t tile = stagger ? (t0i + (int64_t)(blockIdx.x % share) * (gridDim.x / share)) % n_tiles : t0i;
                for (int l = 0; l < loads_per_tile; l++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *st = smem + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                    if (mode == 0) tma_load_2d(st, &map, &full_bar[stage], l * 64, (int)(tile * 64));
                    else if (mode == 1) { for (int k = 0; k < 4; k++) tma_load_2d(st + k * 8192, &map, &full_bar[stage], k * 64, (int)(tile * 64)); }
                    else bulk_load(st, A + tile * 64 * 256, 32768, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (elect_one()) {
        int stage = 0; uint32_t phase = 0;
        for (int64_t t0i = blockIdx.x / share; t0i < n_tiles; t0i += gridDim.x / share) {
            for (int l = 0; l < loads_per_tile; l++) {
                mbar_wait(&full_bar[stage], phase);
                mbar_arrive(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int64_t M = argc > 1 ? atoll(argv[1]) : 1937980;
    __half *A;
    cudaMalloc(&A, (size_t)M * 256 * 2 + 65536);
    cudaMemset(A, 0, (size_t)M * 256 * 2);
    long long *cyc;
    cudaMalloc(&cyc, sizeof(long long) * 256);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    CUtensorMap map;
    cuuint64_t dims[2] = {256, (cuuint64_t)M}; cuuint64_t strides[1] = {512}; cuuint32_t box[2] = {64, 64}; cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const int64_t n_tiles = M / 64;
    constexpr int STAGES = 6;
    cudaFuncSetAttribute(tma_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {42, 144})
      for (int sh = 0; sh < 3; sh++)
        for (int mode = 0; mode < 2; mode++) {
            const int share = sh == 0 ? 1 : 6, stagger = sh == 2;
            const int stage_bytes = mode == 0 ? 8192 : 32768;
            const int64_t nt = grid == 1 ? 2000 : n_tiles;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                tma_kernel<STAGES><<<grid, 64, STAGES * stage_bytes + 1024>>>(map, A, nt, mode, stage_bytes, cyc, share, stagger);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d grid %d: %s\n", mode, grid, cudaGetErrorString(e)); return 1; }
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
            const double bytes = (double)nt * 32768 * share;
            printf("grid %3d share %d stagger %d mode %d (%s, %d KB/stage x %d): %.3f ms  %.1f GB/s delivered  %.1f B/clk/SM\n", grid, share, stagger, mode,
                   mode == 0 ? "tensor 64x64 box" : mode == 1 ? "tensor 4 boxes  " : "linear bulk 32KB", stage_bytes / 1024, STAGES, ms,
                   bytes / ms / 1e6, bytes / grid / (double)mx);
        }
    return 0;
}
