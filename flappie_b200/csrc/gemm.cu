// gemm.cu -- fp32 CUDA-core GEMMs of the hot path (the parity-exact baseline path).
//
//  * sgemm_bias : Xin = A * iW^T + b for all blocks of all reads at once; replaces
//    reference feedforward_linear -> affine_map (src/layers.c:279, src/flappie_matrix.c:361-389,
//    one cblas_sgemm per read per layer).
//  * ff_tanh    : the flip-flop output layer C = tanh(A * FF_W^T + b) / (temperature / 5);
//    replaces the affine_map + tanh_activation_inplace + shift_scale_matrix_inplace of
//    reference globalnorm_manystay (src/layers.c:1082-1087).
//
// The tensor-core (tcgen05) versions live in gemm_tc.cu; these stay as the fp32
// reference path of the library (FFB_FLAG_FP32_SIMT) and for shapes the TC path rejects.
#include "ffb_common.cuh"

namespace ffb {

// ---------------------------------------------------------------------------------
// 128x128x8 register-blocked SGEMM, 256 threads, 8x8 outputs per thread, register
// prefetch of the next k-slab (double buffered shared memory).
// A [M][K] row-major, Wt [K][N] row-major, C [M][N] row-major.  N % 4 == 0, K % 8 == 0.
constexpr int BM = 128, BN = 128, BK = 8;

__global__ void __launch_bounds__(256)
sgemm_bias_kernel(const float *__restrict__ A, const float *__restrict__ Wt, const float *__restrict__ bias,
                  float *__restrict__ C, int64_t M, int N, int K) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t m0 = (int64_t)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    // global -> register staging: A tile 128 rows x 8 k = 256 float4 (along k)
    const int a_row = tid / 2, a_k = (tid % 2) * 4;
    const int b_k = tid / 32, b_n = (tid % 32) * 4;
    const bool a_ok = (m0 + a_row) < M;
    const bool b_ok = (n0 + b_n) < N;   // N % 4 == 0: a float4 is either fully inside or fully outside
    const float *Ap = A + (m0 + a_row) * (int64_t)K + a_k;
    const float *Bp = Wt + (int64_t)b_k * N + n0 + b_n;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

    float4 ra = a_ok ? *reinterpret_cast<const float4 *>(Ap) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 rb = b_ok ? *reinterpret_cast<const float4 *>(Bp) : zero4;
    As[0][a_k + 0][a_row] = ra.x; As[0][a_k + 1][a_row] = ra.y;
    As[0][a_k + 2][a_row] = ra.z; As[0][a_k + 3][a_row] = ra.w;
    *reinterpret_cast<float4 *>(&Bs[0][b_k][b_n]) = rb;
    __syncthreads();

    const int nk = K / BK;
    for (int kt = 0; kt < nk; kt++) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            ra = a_ok ? *reinterpret_cast<const float4 *>(Ap + (kt + 1) * BK) : make_float4(0.f, 0.f, 0.f, 0.f);
            rb = b_ok ? *reinterpret_cast<const float4 *>(Bp + (int64_t)(kt + 1) * BK * N) : zero4;
        }
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4 + 64]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4 + 64]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            const int nxt = cur ^ 1;
            As[nxt][a_k + 0][a_row] = ra.x; As[nxt][a_k + 1][a_row] = ra.y;
            As[nxt][a_k + 2][a_row] = ra.z; As[nxt][a_k + 3][a_row] = ra.w;
            *reinterpret_cast<float4 *>(&Bs[nxt][b_k][b_n]) = rb;
            __syncthreads();
        }
    }

    const bool c0_ok = (n0 + tx * 4) < N, c1_ok = (n0 + tx * 4 + 64) < N;
    const float4 bb0 = c0_ok ? *reinterpret_cast<const float4 *>(bias + n0 + tx * 4) : zero4;
    const float4 bb1 = c1_ok ? *reinterpret_cast<const float4 *>(bias + n0 + tx * 4 + 64) : zero4;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t row = m0 + ty * 4 + (i < 4 ? i : 60 + i);
        if (row >= M) continue;
        float *cp = C + row * (int64_t)N + n0 + tx * 4;
        if (c0_ok)
            *reinterpret_cast<float4 *>(cp) =
                make_float4(acc[i][0] + bb0.x, acc[i][1] + bb0.y, acc[i][2] + bb0.z, acc[i][3] + bb0.w);
        if (c1_ok)
            *reinterpret_cast<float4 *>(cp + 64) =
                make_float4(acc[i][4] + bb1.x, acc[i][5] + bb1.y, acc[i][6] + bb1.z, acc[i][7] + bb1.w);
    }
}

// ---------------------------------------------------------------------------------
// Output layer: N = 40 or 60 columns.  One CTA = 64 rows; the whole Wt (K x N) sits in
// shared memory; thread (row, half) accumulates N/2 outputs.
template <int NH>   // outputs per thread = N / 2
__global__ void __launch_bounds__(128)
ff_tanh_kernel(const float *__restrict__ A, const float *__restrict__ Wt, const float *__restrict__ bias,
               float *__restrict__ C, int64_t M, int K, float scale, int head) {
    constexpr int N = 2 * NH;
    constexpr int KC = 32;   // k-chunk of the A tile
    extern __shared__ __align__(16) float sm[];
    float *Ws = sm;                 // [K][N]
    float *As = sm + (size_t)K * N; // [64][KC + 1]
    const int tid = threadIdx.x;
    for (int i = tid; i < K * N; i += 128) Ws[i] = Wt[i];
    const int r = tid % 64, half = tid / 64;
    const int64_t ntile = (M + 63) / 64;
    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {   // persistent over row tiles
        const int64_t m0 = tile * 64;
        float acc[NH];
#pragma unroll
        for (int j = 0; j < NH; j++) acc[j] = 0.0f;
        for (int k0 = 0; k0 < K; k0 += KC) {
            __syncthreads();
            // 64 rows x 32 k: coalesced 128 B per row
            for (int i = tid; i < 64 * KC; i += 128) {
                const int rr = i / KC, kk = i % KC;
                const int64_t row = m0 + rr;
                As[rr * (KC + 1) + kk] = row < M ? A[row * (int64_t)K + k0 + kk] : 0.0f;
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < KC; kk++) {
                const float a = As[r * (KC + 1) + kk];
                const float *w = Ws + (size_t)(k0 + kk) * N + half * NH;
#pragma unroll
                for (int j = 0; j < NH; j++) acc[j] = fmaf(a, w[j], acc[j]);
            }
        }
        const int64_t row = m0 + r;
        if (row < M) {
            float *cp = C + row * (int64_t)N + half * NH;
#pragma unroll
            for (int j = 0; j < NH; j++) {
                const float v = acc[j] + bias[half * NH + j];
                cp[j] = head ? rle_head(v, half * NH + j, scale) : tanh_ref(v) / scale;   // shift_scale_matrix_inplace: (x - 0) / scale
            }
        }
    }
}

}  // namespace ffb

int ffb_launch_sgemm_bias(const float *A, const float *Wt, const float *bias, float *C, int64_t M, int N, int K,
                          cudaStream_t st) {
    using namespace ffb;
    if (M <= 0) return 0;
    if (N % 4 != 0 || K % BK != 0) return -1;
    const int64_t mt = (M + BM - 1) / BM;
    // gridDim.y is limited to 65535 tiles: split M if needed
    int launches = 0;
    const int64_t max_mt = 65535;
    for (int64_t t0 = 0; t0 < mt; t0 += max_mt) {
        const int64_t nt = (mt - t0) < max_mt ? (mt - t0) : max_mt;
        const int64_t moff = t0 * BM;
        dim3 grid((N + BN - 1) / BN, (unsigned)nt);
        sgemm_bias_kernel<<<grid, 256, 0, st>>>(A + moff * K, Wt, bias, C + moff * N, M - moff < nt * BM ? M - moff : nt * BM, N, K);
        launches++;
    }
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int ffb_launch_ff_tanh(const float *A, const float *Wt, const float *bias, float *C, int64_t M, int N, int K,
                       float scale, int head, cudaStream_t st) {
    using namespace ffb;
    if (M <= 0) return 0;
    const size_t smem = ((size_t)K * N + 64 * 33) * sizeof(float);
    const int64_t ntile = (M + 63) / 64;
    const unsigned grid = (unsigned)(ntile < 148 * 4 ? ntile : 148 * 4);   // <= 4 CTAs per SM, multiple of 148
    if (N == 40) {
        cudaFuncSetAttribute(ff_tanh_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ff_tanh_kernel<20><<<grid, 128, smem, st>>>(A, Wt, bias, C, M, K, scale, head);
    } else if (N == 60) {
        cudaFuncSetAttribute(ff_tanh_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ff_tanh_kernel<30><<<grid, 128, smem, st>>>(A, Wt, bias, C, M, K, scale, head);
    } else {
        return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
