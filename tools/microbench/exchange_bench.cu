// exchange_bench.cu -- how fast can the 8 CTAs of a cluster all-gather a state slice?
// (design input for flappie_b200/csrc/rnn_tc.cu; not part of the library)
//
//   mode 0: cp.async.bulk shared::cta -> shared::cluster, one copy per peer (what rnn_tc does)
//   mode 1: bulk store to an L2-resident global buffer, then ONE multicast bulk load to all 8 CTAs
//   mode 2: st.shared::cluster.v4 from all threads + fence + remote arrive
// Each round: every CTA delivers `bytes` to every CTA of its cluster, waits until all 8 slices have
// landed locally, then tells every peer "consumed" (the rnn_tc handshake).  Prints cycles per round.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_bench exchange_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flappie_b200/csrc/tc_common.cuh"

using namespace ffb::tc;
constexpr int C = 8;

__device__ __forceinline__ void bulk_store_global(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_multicast(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(void *dst_local_alias, uint32_t rank, uint4 v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.v4.b32 [ra], {%2, %3, %4, %5};\n\t}"
        ::"r"(smem_u32(dst_local_alias)), "r"(rank), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// default (.release.cta) remote arrive and default (.acquire.cta) wait, as CUTLASS' ClusterBarrier does
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}

__global__ void __launch_bounds__(256, 1) exch_kernel(int mode, uint32_t bytes, int rounds, uint8_t *gbuf, long long *cycles, uint32_t *check) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full, empty;
    uint8_t *dst = smem;                   // [C][bytes]
    uint8_t *src = smem + (size_t)C * bytes;
    const uint32_t crank = cluster_ctarank();
    const int cluster_id = blockIdx.x / C;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < bytes / 4; i += 256) reinterpret_cast<uint32_t *>(src)[i] = crank * 1000003u + i;
    if (tid == 0) {
        mbar_init(&full, mode == 2 ? C + 1 : 1);
        mbar_init(&empty, C);
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync_all();
    uint8_t *g = gbuf + ((size_t)cluster_id * C + crank) * bytes;
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        const uint32_t ph = r & 1;
        if (mode == 2) {
            if (tid == 0 && r > 0) mbar_wait_cluster(&empty, ph ^ 1u);
            __syncthreads();
            for (uint32_t d = 0; d < C; d++)
                for (uint32_t i = tid; i < bytes / 16; i += 256)
                    st_cluster_v4(dst + crank * bytes + i * 16, d, reinterpret_cast<const uint4 *>(src)[i]);
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                mbar_arrive(&full);   // local "armed"
                for (uint32_t d = 0; d < C; d++) mbar_arrive_remote(&full, d);
                mbar_wait_cluster(&full, ph);
                for (uint32_t d = 0; d < C; d++) mbar_arrive_remote(&empty, d);
            }
        } else if (mode >= 3) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&full, C * bytes);
                if (mode == 4) bulk_store_global(g, src, bytes);
                if (r > 0) mbar_wait(&empty, ph ^ 1u);
                if (mode == 3) {
                    for (uint32_t d = 0; d < C; d++) dsmem_bulk_copy(dst + crank * bytes, src, bytes, &full, d);
                } else {
                    asm volatile("fence.proxy.async;" ::: "memory");
                    bulk_load_multicast(dst + crank * bytes, g, bytes, &full, 0xff);
                }
                mbar_wait(&full, ph);
                for (uint32_t d = 0; d < C; d++) mbar_arrive_remote_relaxed(&empty, d);
            }
        } else if (tid == 0) {
            mbar_arrive_expect_tx(&full, C * bytes);
            if (mode == 1) bulk_store_global(g, src, bytes);
            if (r > 0) mbar_wait_cluster(&empty, ph ^ 1u);
            if (mode == 0) {
                for (uint32_t d = 0; d < C; d++) dsmem_bulk_copy(dst + crank * bytes, src, bytes, &full, d);
            } else {
                asm volatile("fence.proxy.async;" ::: "memory");
                bulk_load_multicast(dst + crank * bytes, g, bytes, &full, 0xff);
            }
            mbar_wait_cluster(&full, ph);
            for (uint32_t d = 0; d < C; d++) mbar_arrive_remote(&empty, d);
        }
    }
    long long t1 = clock64();
    if (tid == 0) {
        mbar_wait_cluster(&empty, (uint32_t)(rounds - 1) & 1u);
        t1 = clock64();
        if (crank == 0) cycles[cluster_id] = (t1 - t0) / rounds;
    }
    __syncthreads();
    cluster_sync_all();
    // checksum of what landed (slice d must hold d*1000003 + i)
    uint32_t bad = 0;
    for (uint32_t d = 0; d < C; d++)
        for (uint32_t i = tid; i < bytes / 4; i += 256)
            bad += reinterpret_cast<uint32_t *>(dst + d * bytes)[i] != d * 1000003u + i;
    if (bad) atomicAdd(check, bad);
}

int main(int argc, char **argv) {
    const int n_clusters = argc > 1 ? atoi(argv[1]) : 13;
    const int rounds = 2000;
    uint8_t *gbuf; long long *cyc; uint32_t *check;
    cudaMalloc(&gbuf, (size_t)n_clusters * C * 65536);
    cudaMalloc(&cyc, sizeof(long long) * n_clusters);
    cudaMalloc(&check, 4);
    cudaFuncSetAttribute(exch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const uint32_t sizes[] = {256, 1024, 2048, 4096, 8192, 10240, 16384};
    for (int mode = 0; mode < 5; mode++)
        for (uint32_t bytes : sizes) {
            cudaMemset(check, 0, 4);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(n_clusters * C); cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = (size_t)(C + 1) * bytes;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, exch_kernel, mode, bytes, rounds, gbuf, cyc, check);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d bytes %u: %s\n", mode, bytes, cudaGetErrorString(e)); return 1; }
            long long h[64]; uint32_t bad;
            cudaMemcpy(h, cyc, sizeof(long long) * n_clusters, cudaMemcpyDeviceToHost);
            cudaMemcpy(&bad, check, 4, cudaMemcpyDeviceToHost);
            long long mx = 0, mn = 1LL << 60;
            for (int i = 0; i < n_clusters; i++) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
            printf("mode %d slice %6u B  (per CTA out/in %7u B)  cycles/round min %6lld max %6lld   in B/cyc %.1f  bad=%u\n", mode, bytes, C * bytes,
                   mn, mx, (double)C * bytes / mx, bad);
        }
    return 0;
}
