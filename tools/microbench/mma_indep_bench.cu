// mma_indep_bench.cu -- tensor-pipe throughput of small-N tcgen05.mma (kind::f16, M=128, A in tensor memory,
// B in shared memory without swizzle) when the MMAs go round-robin into NACC independent accumulators and are
// issued by NISS threads (one per warp) at once, with compile-time operand offsets like the recurrent kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_indep_bench mma_indep_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flappie_b200/csrc/tc_common.cuh"
using namespace ffb::tc;

template <int N, int NACC, int NISS>
__global__ void __launch_bounds__(32 * 8, 1) mma_indep_kernel(int reps, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; i++) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    // warp-uniform role + elect.sync: operands stay in uniform registers (a divergent `lane == 0` branch makes the
    // compiler convert every operand with R2UR, ~45 clk per MMA -- which is all the first version of this bench measured)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp < NISS && elect_one()) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t dB = make_smem_desc(smem_u32(smem), 2 * N * 16, 128, LAYOUT_NONE);
        const uint32_t d = tmem + 256 + warp * NACC * N;     // this issuer's accumulators
        const long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int i = 0; i < 48; i++) {
                const uint64_t ob = (uint64_t)(((i % 16) * 2 * (2 * N * 16)) >> 4);
                umma_f16_ts(d + (i % NACC) * N, tmem + (i % 16) * 8, dB + ob, idesc, 1);
            }
            umma_commit(&bar[warp]);
            mbar_wait(&bar[warp], (uint32_t)r & 1u);
        }
        out[warp] = clock64() - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int NISS>
static void run(long long *d) {
    if (NISS * NACC * N > 256) return;
    cudaFuncSetAttribute(mma_indep_kernel<N, NACC, NISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    const int reps = 256;
    for (int rep = 0; rep < 2; rep++) mma_indep_kernel<N, NACC, NISS><<<1, 256, 70 * 1024>>>(reps, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("N=%3d  %d accumulators/issuer  %d issuers: %7.1f clk per 48-MMA batch+commit+wait per issuer = %5.1f clk/MMA overall (pipe floor %d)\n",
           N, NACC, NISS, (double)h[0] / reps, (double)h[0] / reps / (48.0 * NISS), N / 2);
}

int main() {
    long long *d; cudaMalloc(&d, 64);
    run<16, 1, 1>(d); run<16, 2, 1>(d); run<16, 3, 1>(d); run<16, 4, 1>(d); run<16, 6, 1>(d);
    run<16, 3, 2>(d); run<16, 3, 4>(d); run<16, 3, 5>(d); run<16, 2, 5>(d); run<16, 2, 8>(d);
    run<32, 1, 1>(d); run<32, 2, 1>(d); run<32, 3, 1>(d); run<32, 4, 1>(d);
    run<32, 2, 2>(d); run<32, 2, 4>(d); run<32, 3, 2>(d);
    run<64, 1, 1>(d); run<64, 2, 1>(d); run<64, 3, 1>(d); run<64, 2, 2>(d);
    return 0;
}
