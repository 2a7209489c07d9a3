// signal.cu -- per-read signal preparation on the device: what calculate_post does
// between read_raw and calculate_transitions (reference src/flappie.c:251-259):
//
//   trim_and_segment_raw   src/flappie_common.c:13-28   -> trim_raw_by_mad src/flappie_common.c:47-81
//   medmad_normalise_array src/util.c:198-212           (medianf/madf/quantilef src/util.c:100-187)
//   difference_array + shift_scale_array (--delta)      src/util.c:215-223, 278-287
//
// The reference sorts a copy of the data with qsort for every median.  Here medians are exact order statistics:
//   * chunk MADs (100-sample chunks): one warp per chunk, rank counting in shared memory;
//   * the MAD threshold over a read's chunks: one CTA per read, rank counting;
//   * the read's median and MAD over the trimmed range: one CTA per read, 4-pass radix select on the
//     order-preserving integer image of the floats (the k-th smallest value exactly, no sort), then the
//     (k+1)-th from one more pass; the interpolation between the two is the reference's expression, evaluated
//     in the same mixed float/double arithmetic.
// Everything is integer/compare work on HBM-resident samples: 4 B read per raw sample per pass (L2-resident
// after the first), 4 B written per kept sample.
#include "ffb_common.cuh"

namespace ffb {

// quantile by linear interpolation between the two neighbouring order statistics, src/util.c:125-133:
//   idx = p * (nx - 1); remf = p * (nx - 1) - idx;  (1.0 - remf) * s[idx] + remf * s[idx + 1]
// float * size_t is a float product; `1.0 - remf` and the first product are double, the second product float.
__device__ __forceinline__ void quantile_pos(float p, size_t nx, size_t *idx, float *remf) {
    const float pos = p * (float)(nx - 1);
    *idx = (size_t)pos;
    *remf = pos - (float)(*idx);
}
__device__ __forceinline__ float quantile_mix(float remf, float s0, float s1) {
    return (float)((1.0 - (double)remf) * (double)s0 + (double)(remf * s1));
}

// order-preserving map float -> uint32 (and back)
__device__ __forceinline__ uint32_t fkey(float x) {
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------------------------------------
// warp-wide: the order statistics s[k0] and s[k0+1] (k0 + 1 < n may not hold: then s1 = s0) of v[0..n) in
// shared memory, by rank counting: rank(i) = #{j : v[j] < v[i]} + #{j < i : v[j] == v[i]} is a permutation.
__device__ void warp_two_ranks(const float *v, int n, int k0, float *res /* shared, 2 floats */, int lane) {
    for (int i = lane; i < n; i += 32) {
        const float xi = v[i];
        int r = 0;
        for (int j = 0; j < n; j++) {
            const float xj = v[j];
            r += (xj < xi) || (xj == xi && j < i);
        }
        if (r == k0) res[0] = xi;
        if (r == k0 + 1) res[1] = xi;
    }
    __syncwarp();
}

// MAD of every chunk of every read: madf(raw + start + i*chunk, chunk, NULL) (src/flappie_common.c:58-60).
// One warp per chunk; chunk_off[n] = first chunk of read n in `mad`.
__global__ void __launch_bounds__(256)
chunk_mad_kernel(const float *__restrict__ raw, const int64_t *__restrict__ raw_off, const int64_t *__restrict__ chunk_off,
                 int n_reads, int chunk, int64_t total_chunks, float *__restrict__ mad) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *v = sm + (size_t)warp * (2 * chunk + 2);
    float *a = v + chunk;
    float *res = a + chunk;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= total_chunks) return;
    // read of this chunk: last n with chunk_off[n] <= c
    int lo = 0, hi = n_reads;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= c) lo = mid; else hi = mid;
    }
    const float *x = raw + raw_off[lo] + (c - chunk_off[lo]) * chunk;
    for (int i = lane; i < chunk; i += 32) v[i] = x[i];
    __syncwarp();
    float m = 0.0f;
    if (chunk > 1) {                                   // madf: n == 1 -> 0 (src/util.c:168-170)
        size_t idx; float remf;
        quantile_pos(0.5f, (size_t)chunk, &idx, &remf);
        warp_two_ranks(v, chunk, (int)idx, res, lane);
        const float med = (idx < (size_t)chunk - 1) ? quantile_mix(remf, res[0], res[1]) : res[0];
        for (int i = lane; i < chunk; i += 32) a[i] = fabsf(v[i] - med);
        __syncwarp();
        warp_two_ranks(a, chunk, (int)idx, res, lane);
        const float mabs = (idx < (size_t)chunk - 1) ? quantile_mix(remf, res[0], res[1]) : res[0];
        m = mabs * 1.4826f;                            // mad_scaling_factor, src/util.c:164
    }
    if (lane == 0) mad[c] = m;
}

// trim_raw_by_mad + the trim_start/trim_end arithmetic of trim_and_segment_raw.  One CTA per read.
// bounds[2n] = start, bounds[2n+1] = end (relative to the read); start >= end means "nothing left".
__global__ void __launch_bounds__(256)
trim_bounds_kernel(const float *__restrict__ mad, const int64_t *__restrict__ raw_off, const int64_t *__restrict__ chunk_off,
                   int chunk, float perc, int64_t trim_start, int64_t trim_end, int64_t *__restrict__ bounds) {
    __shared__ float res[2];
    __shared__ int first_hi, last_hi;
    const int n = blockIdx.x, tid = threadIdx.x;
    const int64_t nsample = raw_off[n + 1] - raw_off[n];
    const int nchunk = (int)(chunk_off[n + 1] - chunk_off[n]);
    const float *v = mad + chunk_off[n];
    if (tid == 0) { first_hi = nchunk; last_hi = -1; res[0] = res[1] = 0.0f; }
    __syncthreads();
    int64_t start = 0, end = 0;
    if (nchunk > 0) {
        size_t idx; float remf;
        quantile_pos(perc, (size_t)nchunk, &idx, &remf);
        for (int i = tid; i < nchunk; i += blockDim.x) {
            const float xi = v[i];
            int r = 0;
            for (int j = 0; j < nchunk; j++) {
                const float xj = v[j];
                r += (xj < xi) || (xj == xi && j < i);
            }
            if (r == (int)idx) res[0] = xi;
            if (r == (int)idx + 1) res[1] = xi;
        }
        __syncthreads();
        const float thresh = (idx < (size_t)nchunk - 1) ? quantile_mix(remf, res[0], res[1]) : res[0];
        for (int i = tid; i < nchunk; i += blockDim.x)
            if (v[i] > thresh) { atomicMin(&first_hi, i); atomicMax(&last_hi, i); }
        __syncthreads();
        start = (int64_t)first_hi * chunk;             // chunks before the first one above the threshold
        end = (int64_t)(last_hi + 1) * chunk;          // ... and after the last one
    }
    if (tid == 0) {
        // src/flappie_common.c:19-20 (rt.n = nsample)
        start = (nsample - start) > trim_start ? start + trim_start : nsample;
        end = (end > trim_end) ? end - trim_end : 0;
        bounds[2 * n] = start;
        bounds[2 * n + 1] = end;
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-wide exact order statistics s[k] and s[k+1] of f(x[i]), f = identity or |x - med|.
template <bool ABS>
__device__ __forceinline__ uint32_t sel_key(const float *x, int i, float med) {
    return fkey(ABS ? fabsf(x[i] - med) : x[i]);
}

template <bool ABS>
__device__ void block_two_ranks(const float *x, int n, float med, uint32_t k, uint32_t *hist /* 256 */, uint32_t *scal /* 4 */,
                                float *s0, float *s1) {
    const int tid = threadIdx.x, nt = blockDim.x;
    uint32_t prefix = 0, mask = 0, kk = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = tid; b < 256; b += nt) hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const uint32_t key = sel_key<ABS>(x, i, med);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // warp 0: which bin holds rank kk?  8 bins per lane, then a warp scan
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c[j] = hist[tid * 8 + j]; sum += c[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (tid >= d) incl += o;
            }
            const uint32_t excl = incl - sum;
            if (kk >= excl && kk < incl) {             // exactly one lane
                uint32_t run = excl;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (kk >= run && kk < run + c[j]) { scal[0] = (uint32_t)(tid * 8 + j); scal[1] = kk - run; }
                    run += c[j];
                }
            }
        }
        __syncthreads();
        prefix |= scal[0] << shift;
        mask |= 255u << shift;
        kk = scal[1];
        __syncthreads();
    }
    // prefix = key of s[k].  s[k+1] is the same value if more than k+1 elements are <= it, else the smallest larger one
    if (tid == 0) { scal[2] = 0; scal[3] = 0xffffffffu; }
    __syncthreads();
    uint32_t le = 0, mn = 0xffffffffu;
    for (int i = tid; i < n; i += nt) {
        const uint32_t key = sel_key<ABS>(x, i, med);
        if (key <= prefix) le++; else mn = min(mn, key);
    }
    atomicAdd(&scal[2], le);
    atomicMin(&scal[3], mn);
    __syncthreads();
    *s0 = fkey_inv(prefix);
    *s1 = (scal[2] > k + 1 || scal[3] == 0xffffffffu) ? *s0 : fkey_inv(scal[3]);
    __syncthreads();
}

// One CTA per read: out[sig_off[n] + i] = normalised raw[raw_off[n] + start + i], i < end - start.
// delta == 0: med-MAD normalisation; otherwise difference_array then / delta (src/flappie.c:254-259).
__global__ void __launch_bounds__(512)
normalise_kernel(const float *__restrict__ raw, const int64_t *__restrict__ raw_off, const int64_t *__restrict__ bounds,
                 const int64_t *__restrict__ sig_off, float delta, float *__restrict__ out) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t scal[4];
    const int rd = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int64_t n64 = sig_off[rd + 1] - sig_off[rd];     // 0 for a rejected read
    if (n64 <= 0) return;
    const int n = (int)n64;
    const float *x = raw + raw_off[rd] + bounds[2 * rd];
    float *y = out + sig_off[rd];
    if (delta != 0.0f) {
        for (int i = tid; i < n; i += nt) {
            const float d = (i + 1 < n) ? x[i + 1] - x[i] : 0.0f;
            y[i] = (d - 0.0f) / delta;                      // shift_scale_array(x, n, 0.0, delta)
        }
        return;
    }
    if (n == 1) { if (tid == 0) y[0] = 0.0f; return; }      // src/util.c:202-205
    size_t idx; float remf;
    quantile_pos(0.5f, (size_t)n, &idx, &remf);
    float s0, s1;
    block_two_ranks<false>(x, n, 0.0f, (uint32_t)idx, hist, scal, &s0, &s1);
    const float xmed = (idx < (size_t)n - 1) ? quantile_mix(remf, s0, s1) : s0;
    block_two_ranks<true>(x, n, xmed, (uint32_t)idx, hist, scal, &s0, &s1);
    const float xmad = ((idx < (size_t)n - 1) ? quantile_mix(remf, s0, s1) : s0) * 1.4826f;
    for (int i = tid; i < n; i += nt) y[i] = (x[i] - xmed) / xmad;   // IEEE division, as the reference
}

}  // namespace ffb

int ffb_launch_chunk_mad(const float *raw, const int64_t *raw_off, const int64_t *chunk_off, int n_reads, int chunk,
                         int64_t total_chunks, float *mad, cudaStream_t st) {
    if (total_chunks <= 0) return 0;
    if (chunk < 2 || chunk > FFB_MAX_VARSEG_CHUNK) return -1;
    const int wpb = 8;
    const size_t smem = (size_t)wpb * (2 * chunk + 2) * sizeof(float);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(ffb::chunk_mad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const int64_t grid = (total_chunks + wpb - 1) / wpb;
    ffb::chunk_mad_kernel<<<(unsigned)grid, wpb * 32, smem, st>>>(raw, raw_off, chunk_off, n_reads, chunk, total_chunks, mad);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ffb_launch_trim_bounds(const float *mad, const int64_t *raw_off, const int64_t *chunk_off, int n_reads, int chunk,
                           float perc, int64_t trim_start, int64_t trim_end, int64_t *bounds, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    ffb::trim_bounds_kernel<<<n_reads, 256, 0, st>>>(mad, raw_off, chunk_off, chunk, perc, trim_start, trim_end, bounds);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ffb_launch_normalise(const float *raw, const int64_t *raw_off, const int64_t *bounds, const int64_t *sig_off,
                         int n_reads, float delta, float *out, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    ffb::normalise_kernel<<<n_reads, 512, 0, st>>>(raw, raw_off, bounds, sig_off, delta, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
