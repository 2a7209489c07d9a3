// conv.cu -- strided 1-D convolution + activation over a ragged batch of reads.
//
// Replaces reference features_from_raw (src/nnfeatures.c:15-28), convolution
// (src/layers.c:189-276) and tanh/swish_activation_inplace (src/layers.c:24-48).
//
// Semantics: the reference builds the output from a left-edge, an interior (strided
// SGEMM per phase) and a right-edge part.  For every column except the last few this
// equals the zero-padded "same" convolution; the last columns differ when
// T % stride == 0.  The host replays the reference's integer
// plan per distinct read length and hands the deviating columns over as a ConvTail
// table (<= 2 explicit terms per column); the kernel computes the zero-padded window
// for every other column.
//
// Roofline: HBM-bound on the [blocks][nfilter] fp32 output write.
// One CTA computes a tile of columns of ONE read for all filters: the input span is
// staged once in shared memory (coalesced, zero-filled outside the read), threads map to
// filters so weight reads (Wt[tap*nf+f][filter]) and output writes are coalesced.
#include "ffb_common.cuh"

namespace ffb {

constexpr int CONV_TILE_C = 16;     // output columns per thread
constexpr int CONV_THREADS = 256;

template <int ACT>
__global__ void __launch_bounds__(1024)
conv_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ Wt,
            const float *__restrict__ bias, const ReadGeom *__restrict__ geom,
            const ConvTail *__restrict__ tails, int nf, int nfilter, int winlen, int stride,
            int groups /* column groups per CTA */) {
    extern __shared__ float xs[];   // [span][nf]
    const ReadGeom g = geom[blockIdx.x];
    const int cols_per_cta = groups * CONV_TILE_C;
    const int c0 = blockIdx.y * cols_per_cta;
    if (c0 >= g.T_out) return;
    const int padL = (winlen - 1) / 2;
    const int span = (cols_per_cta - 1) * stride + winlen;   // input columns needed
    const int xin0 = c0 * stride - padL;                     // may be negative
    const float *xr = x + g.in_off * nf;
    for (int i = threadIdx.x; i < span * nf; i += blockDim.x) {
        const int col = xin0 + i / nf;
        xs[i] = (col >= 0 && col < g.T_in) ? xr[(int64_t)col * nf + (i % nf)] : 0.0f;
    }
    __syncthreads();

    const int f = threadIdx.x % nfilter;
    const int grp = threadIdx.x / nfilter;
    if (grp >= groups) return;
    const int cbase = grp * CONV_TILE_C;   // within the CTA tile
    float acc[CONV_TILE_C];
    const float b = bias[f];
#pragma unroll
    for (int c = 0; c < CONV_TILE_C; c++) acc[c] = 0.0f;
    const int K = winlen * nf;
    const float *xsb = xs + cbase * stride * nf;
    for (int j = 0; j < K; j++) {
        const float w = __ldg(Wt + (size_t)j * nfilter + f);
#pragma unroll
        for (int c = 0; c < CONV_TILE_C; c++) acc[c] = fmaf(w, xsb[c * stride * nf + j], acc[c]);
    }
    const ConvTail *tl = tails + g.tail_id;
    const int tail0 = tl->tail_col0;
    float *yr = y + g.out_off * nfilter;
#pragma unroll
    for (int c = 0; c < CONV_TILE_C; c++) {
        const int col = c0 + cbase + c;
        if (col >= g.T_out) break;
        float v = acc[c];
        if (col >= tail0) {
            // explicit terms of the reference plan for the trailing columns
            const int ti = col - tail0;
            v = 0.0f;
            for (int q = 0; q < 2; q++) {
                const int nt = tl->ntap[ti][q];
                if (nt <= 0) continue;
                const float *wq = Wt + (size_t)tl->tap_lo[ti][q] * nf * nfilter + f;
                const float *xq = xr + (int64_t)tl->x_start[ti][q] * nf;
                float a = 0.0f;
                for (int j = 0; j < nt * nf; j++) a = fmaf(__ldg(wq + (size_t)j * nfilter), __ldg(xq + j), a);
                v += a;
            }
        }
        yr[(int64_t)col * nfilter + f] = activate(v + b, ACT);
    }
}

}  // namespace ffb

int ffb_launch_conv(const float *x, float *y, const float *Wt, const float *bias, const ffb::ReadGeom *geom,
                    const ffb::ConvTail *tails, int n_reads, int64_t total_out_cols, int max_T_out, int nf,
                    int nfilter, int winlen, int stride, int act, cudaStream_t st) {
    using namespace ffb;
    (void)total_out_cols;
    if (n_reads <= 0 || max_T_out <= 0) return 0;
    int threads, groups;
    if (nfilter >= CONV_THREADS) {
        threads = nfilter;   // 256 / 384 / 512 filters: one column group
        groups = 1;
        if (threads > 1024) return -1;
    } else {
        groups = CONV_THREADS / nfilter;
        threads = groups * nfilter;
    }
    const int cols_per_cta = groups * CONV_TILE_C;
    const int span = (cols_per_cta - 1) * stride + winlen;
    const size_t smem = (size_t)span * nf * sizeof(float);
    dim3 grid(n_reads, (max_T_out + cols_per_cta - 1) / cols_per_cta);   // x = read, y = column tile
    if (grid.y > 65535) return -1;
    auto launch = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(x, y, Wt, bias, geom, tails, nf, nfilter, winlen, stride, groups);
    };
    if (act == FFB_ACT_TANH) launch(conv_kernel<FFB_ACT_TANH>);
    else if (act == FFB_ACT_SWISH) launch(conv_kernel<FFB_ACT_SWISH>);
    else launch(conv_kernel<FFB_ACT_NONE>);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
