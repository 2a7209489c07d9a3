// cluster_probe.cu -- how many clusters of each size can be co-resident on this chip (one CTA per SM: the
// recurrent kernel allocates all of tensor memory), and which SMs each cluster lands on.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) probe_kernel(int *smid, int spin) {
    extern __shared__ unsigned char smem[];
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) smid[blockIdx.x] = (int)id;
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (spin < 0) smem[threadIdx.x] = 0;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
    int *d; cudaMalloc(&d, 4096 * sizeof(int));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int sizes[] = {1, 2, 4, 6, 7, 8, 10, 12, 14, 16};
    for (int threads : {640, 800, 1024})
        for (int C : sizes) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 120 * 1024;   // > half: one CTA per SM
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
            printf("threads %4d cluster size %2d: max active clusters %3d -> %3d SMs (%s)\n", threads, C, n, n * C, cudaGetErrorString(e));
            cudaGetLastError();
        }
    // where do the CTAs of 15 x 8 / 9 x 16 land?
    for (int C : {8, 16}) {
        cudaLaunchConfig_t cfg = {};
        int n = 0;
        cfg.gridDim = dim3(C * 64); cfg.blockDim = dim3(800); cfg.dynamicSmemBytes = 120 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        if (n <= 0) continue;
        cfg.gridDim = dim3(C * n);
        cudaMemset(d, 0xff, 4096 * sizeof(int));
        cudaLaunchKernelEx(&cfg, probe_kernel, d, 2000000);
        cudaError_t e = cudaDeviceSynchronize();
        int h[4096]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("cluster size %d x %d (%s): SM ids per cluster:\n", C, n, cudaGetErrorString(e));
        for (int c = 0; c < n; c++) { printf("  c%-2d:", c); for (int i = 0; i < C; i++) printf(" %3d", h[c * C + i]); printf("\n"); }
    }
    return 0;
}
