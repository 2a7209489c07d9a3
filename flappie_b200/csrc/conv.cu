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
#include <cuda_fp16.h>
#include <cstdlib>

#include "ffb_common.cuh"

namespace ffb {

constexpr int CONV_TILE_C = 16;     // output columns per thread
constexpr int CONV_THREADS = 256;

__device__ __forceinline__ void conv_store(float *y, __half *yhi, __half *ylo, int64_t idx, int64_t pidx, float v) {
    if (y) y[idx] = v;
    if (yhi) {
        const __half hi = __float2half_rn(v);
        yhi[pidx] = hi;
        ylo[pidx] = __float2half_rn(v - __half2float(hi));
    }
}

// Single-feature input (the raw signal: conv of the GRU topology, first conv of the LSTM topology): the
// thread's whole input window lives in registers, so the inner loop is one weight load per tap and
// CONV_TILE_C register-register FMAs.
// Each thread owns FOUR adjacent filters and CONV1_C columns, so that the fp16 hi/lo planes leave as 8-byte
// stores: with one filter per thread the 2-byte plane stores (two per output) were the kernel's bottleneck on
// the load/store pipe, not the FMAs or the HBM write.
constexpr int CONV1_C = 8;          // output columns per thread
constexpr int CONV1_TILES = 4;      // column tiles one CTA walks through

template <int WINLEN, int STRIDE, int ACT>
__global__ void __launch_bounds__(256, 2)
conv1_kernel(const float *__restrict__ x, float *__restrict__ y, __half *__restrict__ yhi, __half *__restrict__ ylo,
             const float *__restrict__ Wt, const float *__restrict__ bias, const ReadGeom *__restrict__ geom,
             const ConvTail *__restrict__ tails, int nfilter, int groups) {
    extern __shared__ __align__(16) float xs[];   // [WINLEN][nfilter] taps, then [span of CONV1_TILES tiles] samples
    float *ws = xs;
    const ReadGeom g = geom[blockIdx.x];
    const int cols_per_tile = groups * CONV1_C;
    const int cols_per_cta = cols_per_tile * CONV1_TILES;
    const int c0 = blockIdx.y * cols_per_cta;
    if (c0 >= g.T_out) return;
    constexpr int padL = (WINLEN - 1) / 2;
    constexpr int WIN = (CONV1_C - 1) * STRIDE + WINLEN;        // samples one thread needs
    const int span = (cols_per_cta - 1) * STRIDE + WINLEN;
    const int xin0 = c0 * STRIDE - padL;
    const float *xr = x + g.in_off;
    float *xsig = ws + WINLEN * nfilter;
    for (int i = threadIdx.x; i < WINLEN * nfilter; i += blockDim.x) ws[i] = Wt[i];
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const int col = xin0 + i;
        xsig[i] = (col >= 0 && col < g.T_in) ? xr[col] : 0.0f;
    }
    __syncthreads();
    const int fq = nfilter >> 2;
    const int f = 4 * (threadIdx.x % fq);
    const int grp = threadIdx.x / fq;
    if (grp >= groups) return;
    // the thread's four filters: one conflict-free 16-byte shared load per tap
    const float4 b4 = *reinterpret_cast<const float4 *>(bias + f);
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
    const ConvTail *tl = tails + g.tail_id;
    const int tail0 = tl->tail_col0;
#pragma unroll 1
    for (int t = 0; t < CONV1_TILES; t++) {
        const int cbase = t * cols_per_tile + grp * CONV1_C;
        if (c0 + cbase >= g.T_out) break;
        float xw[WIN];
#pragma unroll
        for (int i = 0; i < WIN; i++) xw[i] = xsig[cbase * STRIDE + i];
        float acc[CONV1_C][4];
#pragma unroll
        for (int c = 0; c < CONV1_C; c++) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0f;
#pragma unroll
        for (int j = 0; j < WINLEN; j++) {
            const float4 wj = *reinterpret_cast<const float4 *>(ws + j * nfilter + f);
#pragma unroll
            for (int c = 0; c < CONV1_C; c++) {
                const float xv = xw[c * STRIDE + j];
                acc[c][0] = fmaf(wj.x, xv, acc[c][0]); acc[c][1] = fmaf(wj.y, xv, acc[c][1]);
                acc[c][2] = fmaf(wj.z, xv, acc[c][2]); acc[c][3] = fmaf(wj.w, xv, acc[c][3]);
            }
        }
#pragma unroll
        for (int c = 0; c < CONV1_C; c++) {
            const int col = c0 + cbase + c;
            if (col >= g.T_out) break;
            float v[4] = {acc[c][0], acc[c][1], acc[c][2], acc[c][3]};
            if (col >= tail0) {
                const int ti = col - tail0;
                v[0] = v[1] = v[2] = v[3] = 0.0f;
                for (int q = 0; q < 2; q++) {
                    const int nt = tl->ntap[ti][q];
                    if (nt <= 0) continue;
                    const float *wq = Wt + (size_t)tl->tap_lo[ti][q] * nfilter + f;
                    const float *xq = xr + tl->x_start[ti][q];
                    float a[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < nt; j++) {
                        const float4 wj = __ldg(reinterpret_cast<const float4 *>(wq + (size_t)j * nfilter));
                        const float xj = __ldg(xq + j);
                        a[0] = fmaf(wj.x, xj, a[0]); a[1] = fmaf(wj.y, xj, a[1]); a[2] = fmaf(wj.z, xj, a[2]); a[3] = fmaf(wj.w, xj, a[3]);
                    }
                    v[0] += a[0]; v[1] += a[1]; v[2] += a[2]; v[3] += a[3];
                }
            }
            const int64_t idx = (g.out_off + col) * (int64_t)nfilter + f, pidx = (g.plane_off + col) * (int64_t)nfilter + f;
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = fast_activate(v[k] + bb[k], ACT);
            if (y) *reinterpret_cast<float4 *>(y + idx) = make_float4(o[0], o[1], o[2], o[3]);
            if (yhi) {
                __half h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { h[k] = __float2half_rn(o[k]); l[k] = __float2half_rn(o[k] - __half2float(h[k])); }
                *reinterpret_cast<uint2 *>(yhi + pidx) = *reinterpret_cast<uint2 *>(h);
                *reinterpret_cast<uint2 *>(ylo + pidx) = *reinterpret_cast<uint2 *>(l);
            }
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(1024)
conv_kernel(const float *__restrict__ x, float *__restrict__ y, __half *__restrict__ yhi, __half *__restrict__ ylo, const float *__restrict__ Wt,
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
    if ((nf & 3) == 0) {
        // features are contiguous per input column: four taps per 16-byte shared load
        for (int j = 0; j < K; j += 4) {
            const float w0 = __ldg(Wt + (size_t)j * nfilter + f), w1 = __ldg(Wt + (size_t)(j + 1) * nfilter + f),
                        w2 = __ldg(Wt + (size_t)(j + 2) * nfilter + f), w3 = __ldg(Wt + (size_t)(j + 3) * nfilter + f);
#pragma unroll
            for (int c = 0; c < CONV_TILE_C; c++) {
                const float4 xv = *reinterpret_cast<const float4 *>(xsb + c * stride * nf + j);
                acc[c] = fmaf(w0, xv.x, acc[c]); acc[c] = fmaf(w1, xv.y, acc[c]);
                acc[c] = fmaf(w2, xv.z, acc[c]); acc[c] = fmaf(w3, xv.w, acc[c]);
            }
        }
    } else {
        for (int j = 0; j < K; j++) {
            const float w = __ldg(Wt + (size_t)j * nfilter + f);
#pragma unroll
            for (int c = 0; c < CONV_TILE_C; c++) acc[c] = fmaf(w, xsb[c * stride * nf + j], acc[c]);
        }
    }
    const ConvTail *tl = tails + g.tail_id;
    const int tail0 = tl->tail_col0;
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
        conv_store(y, yhi, ylo, (g.out_off + col) * (int64_t)nfilter + f, (g.plane_off + col) * (int64_t)nfilter + f, fast_activate(v + b, ACT));
    }
}

// Many-feature input, many filters (third convolution of the LSTM topology: 16 -> S, winlen 19, stride 5 -- 90 GMAC per
// 1024 reads, the one convolution that is FMA- and not HBM-bound): register tile of 4 filters x CONV_TILE_C columns per
// thread, so one 16-byte weight load and one 16-byte (broadcast) window load feed 16 FMAs each.
// Requires nf % 4 == 0 and nfilter % 4 == 0; threads = groups * nfilter / 4.
template <int ACT>
__global__ void __launch_bounds__(512)
conv_f4_kernel(const float *__restrict__ x, float *__restrict__ y, __half *__restrict__ yhi, __half *__restrict__ ylo,
               const float *__restrict__ Wt, const float *__restrict__ bias, const ReadGeom *__restrict__ geom,
               const ConvTail *__restrict__ tails, int nf, int nfilter, int winlen, int stride, int groups, int fix_cols) {
    extern __shared__ float xs[];   // [span][nf]
    const ReadGeom g = geom[blockIdx.x];
    const int cols_per_cta = groups * CONV_TILE_C;
    int c0 = blockIdx.y * cols_per_cta, col_end = g.T_out;
    if (fix_cols > 0) {
        // only the head [0, fix_cols) and the tail [T_out - fix_cols, T_out) of the read: half of gridDim.y each
        const int half = gridDim.y >> 1;
        if ((int)blockIdx.y < half) { col_end = min(fix_cols, g.T_out); }
        else { c0 = max(0, g.T_out - fix_cols) + ((int)blockIdx.y - half) * cols_per_cta; }
    }
    if (c0 >= col_end) return;
    const int padL = (winlen - 1) / 2;
    const int span = (cols_per_cta - 1) * stride + winlen;
    const int xin0 = c0 * stride - padL;
    const float *xr = x + g.in_off * nf;
    {
        // nf % 4 == 0: stage whole 16-byte feature quads
        const int nq = nf >> 2;
        for (int i = threadIdx.x; i < span * nq; i += blockDim.x) {
            const int col = xin0 + i / nq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col >= 0 && col < g.T_in) v = *reinterpret_cast<const float4 *>(xr + (int64_t)col * nf + 4 * (i % nq));
            reinterpret_cast<float4 *>(xs)[i] = v;
        }
    }
    __syncthreads();
    const int fq = nfilter >> 2;                       // filter quads
    const int f = 4 * (threadIdx.x % fq);
    const int grp = threadIdx.x / fq;
    if (grp >= groups) return;
    const int cbase = grp * CONV_TILE_C;
    float acc[CONV_TILE_C][4];
#pragma unroll
    for (int c = 0; c < CONV_TILE_C; c++) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0f;
    const int K = winlen * nf;
    const float *xsb = xs + cbase * stride * nf;
    for (int j = 0; j < K; j += 4) {
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)j * nfilter + f));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(j + 1) * nfilter + f));
        const float4 w2 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(j + 2) * nfilter + f));
        const float4 w3 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(j + 3) * nfilter + f));
#pragma unroll
        for (int c = 0; c < CONV_TILE_C; c++) {
            const float4 xv = *reinterpret_cast<const float4 *>(xsb + c * stride * nf + j);
            // per output the taps accumulate in increasing k, like the scalar kernel
            acc[c][0] = fmaf(w0.x, xv.x, acc[c][0]); acc[c][1] = fmaf(w0.y, xv.x, acc[c][1]);
            acc[c][2] = fmaf(w0.z, xv.x, acc[c][2]); acc[c][3] = fmaf(w0.w, xv.x, acc[c][3]);
            acc[c][0] = fmaf(w1.x, xv.y, acc[c][0]); acc[c][1] = fmaf(w1.y, xv.y, acc[c][1]);
            acc[c][2] = fmaf(w1.z, xv.y, acc[c][2]); acc[c][3] = fmaf(w1.w, xv.y, acc[c][3]);
            acc[c][0] = fmaf(w2.x, xv.z, acc[c][0]); acc[c][1] = fmaf(w2.y, xv.z, acc[c][1]);
            acc[c][2] = fmaf(w2.z, xv.z, acc[c][2]); acc[c][3] = fmaf(w2.w, xv.z, acc[c][3]);
            acc[c][0] = fmaf(w3.x, xv.w, acc[c][0]); acc[c][1] = fmaf(w3.y, xv.w, acc[c][1]);
            acc[c][2] = fmaf(w3.z, xv.w, acc[c][2]); acc[c][3] = fmaf(w3.w, xv.w, acc[c][3]);
        }
    }
    const ConvTail *tl = tails + g.tail_id;
    const int tail0 = tl->tail_col0;
    const float4 b4 = *reinterpret_cast<const float4 *>(bias + f);
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int c = 0; c < CONV_TILE_C; c++) {
        const int col = c0 + cbase + c;
        if (col >= col_end) break;
        float v[4] = {acc[c][0], acc[c][1], acc[c][2], acc[c][3]};
        if (col >= tail0) {
            const int ti = col - tail0;
            v[0] = v[1] = v[2] = v[3] = 0.0f;
            for (int q = 0; q < 2; q++) {
                const int nt = tl->ntap[ti][q];
                if (nt <= 0) continue;
                const float *wq = Wt + (size_t)tl->tap_lo[ti][q] * nf * nfilter + f;
                const float *xq = xr + (int64_t)tl->x_start[ti][q] * nf;
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < nt * nf; j++) {
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(wq + (size_t)j * nfilter));
                    const float xj = __ldg(xq + j);
                    a[0] = fmaf(w.x, xj, a[0]); a[1] = fmaf(w.y, xj, a[1]); a[2] = fmaf(w.z, xj, a[2]); a[3] = fmaf(w.w, xj, a[3]);
                }
                v[0] += a[0]; v[1] += a[1]; v[2] += a[2]; v[3] += a[3];
            }
        }
        const int64_t idx = (g.out_off + col) * (int64_t)nfilter + f, pidx = (g.plane_off + col) * (int64_t)nfilter + f;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = fast_activate(v[k] + bb[k], ACT);
        if (y) *reinterpret_cast<float4 *>(y + idx) = make_float4(o[0], o[1], o[2], o[3]);
        if (yhi) {
            __half h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { h[k] = __float2half_rn(o[k]); l[k] = __float2half_rn(o[k] - __half2float(h[k])); }
            *reinterpret_cast<uint2 *>(yhi + pidx) = *reinterpret_cast<uint2 *>(h);
            *reinterpret_cast<uint2 *>(ylo + pidx) = *reinterpret_cast<uint2 *>(l);
        }
    }
}

}  // namespace ffb

int ffb_launch_conv(const float *x, float *y, void *yhi_, void *ylo_, const float *Wt, const float *bias,
                    const ffb::ReadGeom *geom, const ffb::ConvTail *tails, int n_reads, int64_t total_out_cols,
                    int max_T_out, int nf, int nfilter, int winlen, int stride, int act, int fix_cols, cudaStream_t st) {
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
    __half *yhi = (__half *)yhi_, *ylo = (__half *)ylo_;
    const int cols_per_cta = groups * CONV_TILE_C;
    const int span = (cols_per_cta - 1) * stride + winlen;
    const size_t smem = (size_t)span * nf * sizeof(float);
    dim3 grid(n_reads, (max_T_out + cols_per_cta - 1) / cols_per_cta);   // x = read, y = column tile
    if (grid.y > 65535) return -1;
    if (nf == 1 && (nfilter & 3) == 0 && nfilter <= 1024 && ((winlen == 19 && stride == 2) || (winlen == 5 && stride == 1))) {
        // four filters x CONV1_C columns per thread
        const int fq = nfilter / 4;
        const int groups1 = fq >= 256 ? 1 : 256 / fq;
        const int threads1 = groups1 * fq;
        const int cols1 = groups1 * CONV1_C * CONV1_TILES;
        const size_t smem1 = (size_t)((cols1 - 1) * stride + winlen + winlen * nfilter) * sizeof(float);
        dim3 grid1(n_reads, (max_T_out + cols1 - 1) / cols1);
        if (grid1.y > 65535 || smem1 > 200 * 1024) return -1;
        auto launch1 = [&](auto kern) {
            if (smem1 > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
            kern<<<grid1, threads1, smem1, st>>>(x, y, yhi, ylo, Wt, bias, geom, tails, nfilter, groups1);
        };
        if (winlen == 19) {
            if (act == FFB_ACT_TANH) launch1(conv1_kernel<19, 2, FFB_ACT_TANH>);
            else if (act == FFB_ACT_SWISH) launch1(conv1_kernel<19, 2, FFB_ACT_SWISH>);
            else launch1(conv1_kernel<19, 2, FFB_ACT_NONE>);
        } else {
            if (act == FFB_ACT_TANH) launch1(conv1_kernel<5, 1, FFB_ACT_TANH>);
            else if (act == FFB_ACT_SWISH) launch1(conv1_kernel<5, 1, FFB_ACT_SWISH>);
            else launch1(conv1_kernel<5, 1, FFB_ACT_NONE>);
        }
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    if (fix_cols > 0 && !((nf & 3) == 0 && (nfilter & 3) == 0 && nfilter >= 128 && nfilter / 4 <= 512)) return -1;
    if ((nf & 3) == 0 && (nfilter & 3) == 0 && nfilter >= 128 && nfilter / 4 <= 512 && (fix_cols > 0 || getenv("FFB_CONV_SCALAR") == nullptr)) {
        // 4 filters per thread; two or three column groups per CTA (one for the narrow fix-up launches)
        const int fq = nfilter / 4;
        const int groups4 = fix_cols > 0 ? 1 : (fq >= 256 ? 1 : (fq >= 128 ? 2 : 3));
        const int threads4 = groups4 * fq;
        const int cols4 = groups4 * CONV_TILE_C;
        const size_t smem4 = (size_t)((cols4 - 1) * stride + winlen) * nf * sizeof(float);
        dim3 grid4(n_reads, fix_cols > 0 ? 2 * ((fix_cols + cols4 - 1) / cols4) : (max_T_out + cols4 - 1) / cols4);
        if (grid4.y <= 65535) {
            auto launch4 = [&](auto kern) {
                if (smem4 > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
                kern<<<grid4, threads4, smem4, st>>>(x, y, yhi, ylo, Wt, bias, geom, tails, nf, nfilter, winlen, stride, groups4, fix_cols);
            };
            if (act == FFB_ACT_TANH) launch4(conv_f4_kernel<FFB_ACT_TANH>);
            else if (act == FFB_ACT_SWISH) launch4(conv_f4_kernel<FFB_ACT_SWISH>);
            else launch4(conv_f4_kernel<FFB_ACT_NONE>);
            return cudaGetLastError() == cudaSuccess ? 1 : -1;
        }
    }
    auto launch = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, st>>>(x, y, yhi, ylo, Wt, bias, geom, tails, nf, nfilter, winlen, stride, groups);
    };
    if (act == FFB_ACT_TANH) launch(conv_kernel<FFB_ACT_TANH>);
    else if (act == FFB_ACT_SWISH) launch(conv_kernel<FFB_ACT_SWISH>);
    else launch(conv_kernel<FFB_ACT_NONE>);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
