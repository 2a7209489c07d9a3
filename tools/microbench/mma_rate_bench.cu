// mma_rate_bench.cu -- cycles per tcgen05.mma (kind::f16, M=128) for the shapes the library uses:
// A from tensor memory or shared memory, B from shared memory (no swizzle / 128-byte swizzle), N = 16..256.
// Operands are garbage (zeros); only the issue/retire rate matters.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_bench mma_rate_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flappie_b200/csrc/tc_common.cuh"
using namespace ffb::tc;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int a_in_tmem, int swz, int n_mma, int n_acc, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 32) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem) + 16384;
        uint64_t db[4], da[4];
        for (int ks = 0; ks < 4; ks++) {
            db[ks] = swz ? make_smem_desc(b_s + ks * 32, 16, 1024, LAYOUT_SW128) : make_smem_desc(b_s + ks * 2 * N * 16, N * 16, 128, LAYOUT_NONE);
            da[ks] = swz ? make_smem_desc(a_s + ks * 32, 16, 1024, LAYOUT_SW128) : make_smem_desc(a_s + ks * 2 * 128 * 16, 128 * 16, 128, LAYOUT_NONE);
        }
        const uint32_t d = tmem + 256;
        const long long t0 = clock64();
        if (a_in_tmem) {
            for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) umma_f16_ts(d + (n_acc > 1 ? ks * 0 : 0), tmem + ks * 8, db[ks], idesc, 1);
            }
        } else {
            for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) umma_f16(d, da[ks], db[ks], idesc, 1);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[0] = clock64() - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int n_mma = 4096;
    for (int a_in_tmem = 0; a_in_tmem < 2; a_in_tmem++)
        for (int swz = 0; swz < 2; swz++)
            for (int N : {16, 32, 64, 128, 256}) {
                if (!swz && N * 16 * 8 > 32768) continue;
                for (int rep = 0; rep < 2; rep++) mma_rate_kernel<<<1, 128, 50 * 1024>>>(N, a_in_tmem, swz, n_mma, 1, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("A in %s, B %s, N=%3d: %.1f clk/MMA (pipe floor %d)\n", a_in_tmem ? "TMEM" : "smem", swz ? "SW128 " : "noswz", N, (double)h / n_mma, N / 2);
            }
    return 0;
}
