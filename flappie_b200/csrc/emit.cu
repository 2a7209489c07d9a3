// emit.cu -- base / quality emission of calculate_post on the device.
//
// Replaces reference change_positions (src/decode.c:66-79) + the emission loop of calculate_post
// (src/flappie.c:284-297) + phredf / qscoref (src/util.h:285-305) + reverse_char_array (src/util.c:416):
//   a base is emitted at every pos in [1, nblock) with path[pos] != path[pos-1]:  "ACGTZ"[path[pos] % nbase],
//   quality char = round(33 - 10 log10(1 - min(p, 0.99999))), p = expf(qpath[pos]).
// One warp per read: a counting pass, then a ballot / prefix compaction that writes the characters (mirrored when
// --reverse).  D2H per batch drops from path + qpath (8 B per block) to 2 B per block.
//
// Bit-exactness against the host's libm: the quality character is a non-decreasing step function of qpath with ~50
// steps.  The host computes -- with its own expf / log1pf, exactly as ffb_emit_bases does -- the smallest float at
// which each character is reached (bisection over the float ordering, api.cu:phred_thresholds) and the device only
// compares against that table, so no device transcendental is involved.
#include "ffb_common.cuh"

namespace ffb {

__global__ void __launch_bounds__(128) emit_kernel(const int32_t *__restrict__ path, const float *__restrict__ qpath,
                                                   const int64_t *__restrict__ blk_off, int n_reads, int nbase, int reverse,
                                                   const float *__restrict__ thr, int nthr, char *__restrict__ bases,
                                                   char *__restrict__ quals, int32_t *__restrict__ nbases) {
    __shared__ float s_thr[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_thr[i] = i < nthr ? thr[i] : __int_as_float(0x7f800000);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int rd = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (rd >= n_reads) return;
    const int64_t b0 = blk_off[rd];
    const int T = (int)(blk_off[rd + 1] - b0);
    const int64_t o0 = b0 + rd;                         // read n owns entries [blk_off[n] + n, blk_off[n+1] + n + 1)
    if (T <= 0) { if (lane == 0) nbases[rd] = 0; return; }
    const int32_t *p = path + o0;
    const float *q = qpath + o0;
    // ---- pass 1: how many bases ----
    int total = 0;
    for (int pos0 = 1; pos0 < T; pos0 += 32) {
        const int pos = pos0 + lane;
        const bool f = pos < T && p[pos] != p[pos - 1];
        total += __popc(__ballot_sync(0xffffffffu, f));
    }
    // ---- pass 2: write them ----
    char *bo = bases + o0, *qo = quals + o0;
    int done = 0;
    for (int pos0 = 1; pos0 < T; pos0 += 32) {
        const int pos = pos0 + lane;
        int st = 0;
        bool f = false;
        if (pos < T) { st = p[pos]; f = st != p[pos - 1]; }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) {
            const int idx = done + __popc(m & ((1u << lane) - 1u));
            const int o = reverse ? total - 1 - idx : idx;
            const float x = q[pos];
            // number of thresholds <= x (upper bound in an ascending table); NaN compares false everywhere: the host's
            // `p < 0.99999` is false for NaN too and clips to the top character
            int lo = 0, hi = nthr;
            if (x != x) lo = nthr;
            else while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_thr[mid] <= x) lo = mid + 1; else hi = mid; }
            bo[o] = "ACGTZ"[st % nbase];
            qo[o] = (char)(33 + lo);
        }
        done += __popc(m);
    }
    if (lane == 0) { bo[total] = 0; qo[total] = 0; nbases[rd] = total; }
}

}  // namespace ffb

int ffb_launch_emit(const int32_t *path, const float *qpath, const int64_t *blk_off, int n_reads, int nbase, int reverse,
                    const float *thr, int nthr, char *bases, char *quals, int32_t *nbases, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    if (nbase < 1 || nbase > 5 || nthr < 0 || nthr > 128) return -1;
    ffb::emit_kernel<<<(n_reads + 3) / 4, 128, 0, st>>>(path, qpath, blk_off, n_reads, nbase, reverse, thr, nthr, bases, quals, nbases);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
