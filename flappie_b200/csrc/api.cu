// api.cu -- host side of libflappie_b200.so: the C ABI declared in include/flappie_b200.h.
//
// Holds (a) the model arena (weights unpacked from the reference's `_Mat` bundle and
// re-laid-out for the kernels), (b) the per-stream context with grow-only HBM workspaces,
// (c) the host planning of a ragged batch (block counts, length sort, the reference's
// convolution edge plan) and (d) the per-read drop-ins with the reference's signatures.
//
// There is deliberately NO CPU implementation of any arithmetic in this file: if CUDA is
// unavailable every entry point fails loudly (NULL / NAN / negative status + message).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/flappie_b200.h"
#include "ffb_common.cuh"

// ------------------------------------------------------------------------------------
static thread_local std::string g_err;
static void set_err(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    fprintf(stderr, "flappie_b200: %s\n", buf);
}
#define CUDA_TRY(expr, ret)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            set_err("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return ret;                                                                  \
        }                                                                                \
    } while (0)

extern "C" const char *ffb_last_error(void) { return g_err.c_str(); }
extern "C" const char *ffb_version(void) { return "flappie_b200 0.1 (sm_100a)"; }
extern "C" int ffb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------
// Grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
            cap = bytes;
            return 0;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Grow-only PINNED host staging area for the small per-batch plan arrays (offsets, geometry, schedules): copies out of
// pageable memory make cudaMemcpyAsync wait for the stream first, copies out of this do not, and its contents live until
// the context's next batch -- so nothing has to be synchronised just to keep a std::vector alive.
struct PinArena {
    uint8_t *p = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes) {            // invalidates earlier contents: callers reserve once per batch phase, up front
        used = 0;
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc((void **)&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
        cap = want;
        return 0;
    }
    void *take(size_t bytes) {             // 16-byte aligned slice; the caller has reserved enough
        const size_t at = (used + 15) & ~(size_t)15;
        if (at + bytes > cap) return nullptr;
        used = at + bytes;
        return p + at;
    }
    void *put(const void *src, size_t bytes) {
        void *d = take(bytes);
        if (d && bytes) memcpy(d, src, bytes);
        return d;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = used = 0; }
};

// ------------------------------------------------------------------------------------
struct ffb_model {
    int device = 0, kind = 0, S = 0, G = 0, nparam = 0, nbase = 0, nstate = 0, nconv = 0;
    // last convolution of the LSTM topology on the tensor cores: W [nfilter][Kp] fp16 planes (K = winlen * nf, zero-padded to Kp)
    void *d_c3_hi = nullptr, *d_c3_lo = nullptr;
    int c3_kp = 0;
    bool tc_conv3 = false;
    bool simt_rnn = true;   // the fp32 CUDA-core recurrence exists for this size
    int head = 0;     // 0 = flip-flop CRF, 1 = run-length CRF (FFB_KIND_RUNLENGTH: LSTM topology, runnie head)
    int conv_nf[FFB_MAX_CONV] = {0}, conv_nfilter[FFB_MAX_CONV] = {0}, conv_winlen[FFB_MAX_CONV] = {0},
        conv_stride[FFB_MAX_CONV] = {0};
    float *d_convWt[FFB_MAX_CONV] = {nullptr}, *d_convb[FFB_MAX_CONV] = {nullptr};
    float *d_iWt[FFB_NLAYER] = {nullptr}, *d_b[FFB_NLAYER] = {nullptr}, *d_sWp[FFB_NLAYER] = {nullptr};
    float *d_ffWt = nullptr, *d_ffb = nullptr;
    void *d_ff_hi = nullptr, *d_ff_lo = nullptr;   // fp16 planes of FF_W [FFB_FF_TC_ROWS][S], zero-padded (tensor output layer)
    float *d_ffb_pad = nullptr;
    bool tc_ff = false;
    float *d_phred_thr = nullptr;   // emit.cu: qpath thresholds of the quality characters
    int n_phred_thr = 0;
    void *d_iW_hi[FFB_NLAYER] = {nullptr}, *d_iW_lo[FFB_NLAYER] = {nullptr};   // fp16 planes [G*S][in] for the tensor path
    bool tc_gemm = false;
    void *d_sW_img[FFB_NLAYER] = {nullptr};   // per-CTA shared-memory images of sW (fp16 hi/lo) for rnn_tc
    bool tc_rnn = false;
    bool fuse_z = false;      // GRU: recurrent layer l also computes the z-gate third of layer l+1's input projection (rnn_tc.cu)
    bool fuse_ff = false;     // ... and the top layer the flip-flop output layer (FF_W rows in the same free quarter)
    float *d_ffb_s = nullptr; // FF bias zero-padded to S entries
    int tc_max_clusters = 0;
    int layer_in[FFB_NLAYER] = {0};
    // Head gate (pipelined contexts of one model): the context whose batch most recently ENTERED its last recurrent layer,
    // and the event it recorded there.  The next batch of another context starts its device work (signal preparation,
    // convolution, first input projection) behind that event: the last recurrent layer has no streamed GEMM beside it,
    // so its spare SMs are idle, whereas earlier in the step every SM is taken and the new batch's head kernels would
    // only push the batch in flight back by their own duration.
    std::mutex gate_mu;
    cudaEvent_t gate_ev = nullptr;
    const struct ffb_ctx *gate_owner = nullptr;
    // conv edge plans, cached per (conv layer, T_in)
    std::mutex mu;
    std::map<std::pair<int, int>, ffb::ConvTail> tail_cache;   // at most FFB_TAIL_CACHE_MAX entries (~0.8 KB each)
};

static constexpr size_t FFB_TAIL_CACHE_MAX = 8192;
static const std::vector<float> &phred_thresholds();   // quality-character steps for the device emission (below)
static inline float mat_at(const _Mat *m, size_t r, size_t c) { return m->data.f[c * m->stride + r]; }

static float *upload(const std::vector<float> &h) {
    float *d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); cudaFree(d); return nullptr;
    }
    return d;
}

extern "C" void ffb_model_destroy(ffb_model *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (int i = 0; i < FFB_MAX_CONV; i++) { cudaFree(m->d_convWt[i]); cudaFree(m->d_convb[i]); }
    for (int i = 0; i < FFB_NLAYER; i++) { cudaFree(m->d_iWt[i]); cudaFree(m->d_b[i]); cudaFree(m->d_sWp[i]); cudaFree(m->d_iW_hi[i]); cudaFree(m->d_iW_lo[i]); cudaFree(m->d_sW_img[i]); }
    cudaFree(m->d_ffWt); cudaFree(m->d_ffb);
    cudaFree(m->d_ff_hi); cudaFree(m->d_ff_lo); cudaFree(m->d_ffb_pad);
    cudaFree(m->d_c3_hi); cudaFree(m->d_c3_lo);
    cudaFree(m->d_phred_thr); cudaFree(m->d_ffb_s);
    delete m;
}

extern "C" ffb_model *ffb_model_create(int device, int kind, const _Mat *const *mats, int nmat,
                                       const int *conv_stride, int nconv) {
    if (!mats || !conv_stride) { set_err("ffb_model_create: NULL argument"); return nullptr; }
    if (ffb_device_count() <= device) { set_err("ffb_model_create: CUDA device %d not available", device); return nullptr; }
    const int head = (kind == FFB_KIND_RUNLENGTH) ? 1 : 0;
    if (head) kind = FFB_KIND_LSTM;                      // runlength5_guppy_transitions: the LSTM stack (networks.c:675-722)
    const int want_conv = (kind == FFB_KIND_GRU) ? 1 : 3;
    if ((kind != FFB_KIND_GRU && kind != FFB_KIND_LSTM) || nconv != want_conv || nmat != 2 * nconv + 3 * FFB_NLAYER + 2) {
        set_err("ffb_model_create: kind %d expects %d convolutions and %d matrices (got %d, %d)", kind, want_conv,
                2 * want_conv + 17, nconv, nmat);
        return nullptr;
    }
    for (int i = 0; i < nmat; i++)
        if (!mats[i] || !mats[i]->data.f) { set_err("ffb_model_create: matrix %d is NULL", i); return nullptr; }
    CUDA_TRY(cudaSetDevice(device), nullptr);
    ffb_model *m = new ffb_model();
    m->device = device; m->kind = kind; m->nconv = nconv; m->head = head;
    m->G = (kind == FFB_KIND_GRU) ? 3 : 4;
    bool ok = true;
    int nf = 1;
    for (int i = 0; i < nconv; i++) {
        const _Mat *W = mats[2 * i], *b = mats[2 * i + 1];
        const int nf4 = 4 * ((nf + 3) / 4);
        const int nfilter = (int)W->nc;
        const int winlen = (int)((W->nr - nf + nf4) / nf4);   // nr = nf4*winlen - nf4 + nf
        if ((size_t)(nf4 * winlen - nf4 + nf) != W->nr || b->nr != W->nc || conv_stride[i] < 1) {
            set_err("ffb_model_create: convolution %d has inconsistent shape (nr=%zu nc=%zu nf=%d)", i, W->nr, W->nc, nf);
            ok = false; break;
        }
        m->conv_nf[i] = nf; m->conv_nfilter[i] = nfilter; m->conv_winlen[i] = winlen; m->conv_stride[i] = conv_stride[i];
        std::vector<float> Wt((size_t)winlen * nf * nfilter), bb(nfilter);
        for (int f = 0; f < nfilter; f++) {
            for (int k = 0; k < winlen; k++)
                for (int n = 0; n < nf; n++) Wt[((size_t)k * nf + n) * nfilter + f] = mat_at(W, (size_t)k * nf4 + n, f);
            bb[f] = mat_at(b, f, 0);
        }
        m->d_convWt[i] = upload(Wt); m->d_convb[i] = upload(bb);
        ok = ok && m->d_convWt[i] && m->d_convb[i];
        if (ok && nconv == 3 && i == 2) {
            // im2col weight planes [filter][k = tap * nf + feature], zero beyond winlen * nf
            const int K = winlen * nf, Kp = 64 * ((K + 63) / 64);
            if (ffb_conv_tc_supported(nfilter, Kp) && (nf * conv_stride[i] * 2) % 16 == 0 && getenv("FFB_NO_TC_CONV") == nullptr) {
                std::vector<float> Wd((size_t)nfilter * Kp, 0.0f);
                for (int f = 0; f < nfilter; f++)
                    for (int k = 0; k < K; k++) Wd[(size_t)f * Kp + k] = Wt[(size_t)k * nfilter + f];
                float *tmp = upload(Wd);
                const bool okc = tmp && cudaMalloc(&m->d_c3_hi, Wd.size() * 2) == cudaSuccess && cudaMalloc(&m->d_c3_lo, Wd.size() * 2) == cudaSuccess &&
                                 ffb_launch_split_f16(tmp, m->d_c3_hi, m->d_c3_lo, (int64_t)Wd.size(), 0) >= 0 && cudaDeviceSynchronize() == cudaSuccess;
                cudaFree(tmp);
                if (okc) { m->tc_conv3 = true; m->c3_kp = Kp; }
                else cudaGetLastError();
            }
        }
        nf = nfilter;
    }
    const _Mat *const *L = mats + 2 * nconv;
    if (ok) {
        m->S = (int)L[1]->nr;   // sW is [S x G*S]
        const int S = m->S, G = m->G;
        if (!ffb_rnn_supported(kind, S) && !ffb_rnn_tc_supported(kind, S)) {
            set_err("ffb_model_create: no recurrent kernel for kind %d size %d", kind, S);
            ok = false;
        }
        m->simt_rnn = ffb_rnn_supported(kind, S) != 0;    // every shipped size (S = 512: weights half-resident, cross-check speed)
        const bool fuse_z = ffb_rnn_tc_can_fuse_z(kind, S) && ffb_gemm_tc_stream_supported(G * S, S) && getenv("FFB_NO_FUSE_Z") == nullptr;
        m->fuse_z = fuse_z;
        int in = nf;
        for (int l = 0; ok && l < FFB_NLAYER; l++) {
            const _Mat *iW = L[3 * l], *sW = L[3 * l + 1], *b = L[3 * l + 2];
            if ((int)iW->nr != in || (int)iW->nc != G * S || (int)sW->nr != S || (int)sW->nc != G * S || (int)b->nr != G * S) {
                set_err("ffb_model_create: recurrent layer %d has inconsistent shape", l);
                ok = false; break;
            }
            m->layer_in[l] = in;
            std::vector<float> iWt((size_t)in * G * S), bb(G * S), sWd((size_t)G * S * S), packed(m->simt_rnn ? ffb_rnn_packed_floats(kind, S) : 1);
            for (int n = 0; n < G * S; n++) {
                for (int k = 0; k < in; k++) iWt[(size_t)k * G * S + n] = mat_at(iW, k, n);
                for (int k = 0; k < S; k++) sWd[(size_t)n * S + k] = mat_at(sW, k, n);
                bb[n] = mat_at(b, n, 0);
            }
            if (m->simt_rnn) ffb_rnn_pack_weights(kind, S, sWd.data(), packed.data());
            m->d_iWt[l] = upload(iWt); m->d_b[l] = upload(bb); m->d_sWp[l] = upload(packed);
            ok = ok && m->d_iWt[l] && m->d_b[l] && m->d_sWp[l];
            if (ok && ffb_rnn_tc_supported(kind, S)) {
                // the free quarter of a GRU layer's M=128 tiles carries the z-gate rows of the NEXT layer's projection
                std::vector<float> zrows;
                if (fuse_z && l + 1 == FFB_NLAYER && !head && getenv("FFB_NO_FUSE_FF") == nullptr) {
                    // top layer: the flip-flop output layer's rows (FF_W is [S x nparam], nparam <= S), zero beyond
                    const _Mat *FW = L[15];
                    if ((int)FW->nr == S && (int)FW->nc <= S) {
                        zrows.assign((size_t)S * S, 0.0f);
                        for (int n = 0; n < (int)FW->nc; n++)
                            for (int k = 0; k < S; k++) zrows[(size_t)n * S + k] = mat_at(FW, k, n);
                        m->fuse_ff = true;
                    }
                }
                if (fuse_z && l + 1 < FFB_NLAYER) {
                    const _Mat *iWn = L[3 * (l + 1)];
                    if ((int)iWn->nr == S && (int)iWn->nc == G * S) {
                        zrows.resize((size_t)S * S);
                        for (int n = 0; n < S; n++)
                            for (int k = 0; k < S; k++) zrows[(size_t)n * S + k] = mat_at(iWn, k, n);
                    }
                }
                std::vector<uint16_t> img(ffb_rnn_tc_image_halfs(kind, S));
                ffb_rnn_tc_pack(kind, S, sWd.data(), img.data(), zrows.empty() ? nullptr : zrows.data());
                ok = cudaMalloc(&m->d_sW_img[l], img.size() * 2) == cudaSuccess &&
                     cudaMemcpy(m->d_sW_img[l], img.data(), img.size() * 2, cudaMemcpyHostToDevice) == cudaSuccess;
            }
            if (ok && ffb_gemm_tc_supported(G * S, in)) {
                // hi/lo fp16 planes of iW in the reference's own [out][in] orientation (K-major B operand)
                std::vector<float> iWd((size_t)G * S * in);
                for (int n = 0; n < G * S; n++)
                    for (int k = 0; k < in; k++) iWd[(size_t)n * in + k] = mat_at(iW, k, n);
                float *tmp = upload(iWd);
                ok = tmp && cudaMalloc(&m->d_iW_hi[l], iWd.size() * 2) == cudaSuccess && cudaMalloc(&m->d_iW_lo[l], iWd.size() * 2) == cudaSuccess &&
                     ffb_launch_split_f16(tmp, m->d_iW_hi[l], m->d_iW_lo[l], (int64_t)iWd.size(), 0) >= 0 && cudaDeviceSynchronize() == cudaSuccess;
                cudaFree(tmp);
            }
            in = S;
        }
    }
    if (ok) {
        const _Mat *FW = L[15], *Fb = L[16];
        m->nparam = (int)FW->nc;
        m->nbase = (int)nbase_from_flipflop_nparam(m->nparam);
        m->nstate = 2 * m->nbase;
        if ((int)FW->nr != m->S || (int)Fb->nr != m->nparam || m->nstate * (m->nbase + 1) != m->nparam ||
            (m->nparam != 40 && m->nparam != 60) || (m->head && m->nparam != 40)) {
            set_err("ffb_model_create: output layer has unsupported shape (%zu x %zu)", FW->nr, FW->nc);
            ok = false;
        } else {
            std::vector<float> Wt((size_t)m->S * m->nparam), bb(m->nparam);
            for (int n = 0; n < m->nparam; n++) {
                for (int k = 0; k < m->S; k++) Wt[(size_t)k * m->nparam + n] = mat_at(FW, k, n);
                bb[n] = mat_at(Fb, n, 0);
            }
            m->d_ffWt = upload(Wt); m->d_ffb = upload(bb);
            ok = m->d_ffWt && m->d_ffb;
            if (ok && m->fuse_ff) {
                std::vector<float> bs((size_t)m->S, 0.0f);
                std::copy(bb.begin(), bb.end(), bs.begin());
                m->d_ffb_s = upload(bs);
                ok = m->d_ffb_s != nullptr;
            }
            if (ok && ffb_ff_tc_supported(m->nparam, m->S)) {
                // [out][in] planes like iW, zero rows beyond nparam
                std::vector<float> Wd((size_t)FFB_FF_TC_ROWS * m->S, 0.0f), bp(FFB_FF_TC_ROWS, 0.0f);
                for (int n = 0; n < m->nparam; n++) {
                    for (int k = 0; k < m->S; k++) Wd[(size_t)n * m->S + k] = mat_at(FW, k, n);
                    bp[n] = bb[n];
                }
                float *tmp = upload(Wd);
                m->d_ffb_pad = upload(bp);
                ok = tmp && m->d_ffb_pad && cudaMalloc(&m->d_ff_hi, Wd.size() * 2) == cudaSuccess &&
                     cudaMalloc(&m->d_ff_lo, Wd.size() * 2) == cudaSuccess &&
                     ffb_launch_split_f16(tmp, m->d_ff_hi, m->d_ff_lo, (int64_t)Wd.size(), 0) >= 0 && cudaDeviceSynchronize() == cudaSuccess;
                cudaFree(tmp);
                m->tc_ff = ok;
            }
        }
    }
    if (ok) {
        m->tc_gemm = true;
        for (int l = 0; l < FFB_NLAYER; l++) m->tc_gemm = m->tc_gemm && m->d_iW_hi[l] && m->d_iW_lo[l];
        m->d_phred_thr = upload(phred_thresholds());
        m->n_phred_thr = (int)phred_thresholds().size();
        ok = m->d_phred_thr != nullptr;
    }
    if (ok && ffb_rnn_tc_supported(kind, m->S) && m->tc_gemm) {
        if (ffb_rnn_tc_prepare(kind, m->S) == 0) {
            m->tc_max_clusters = ffb_rnn_tc_max_clusters(kind, m->S, ffb_rnn_tc_rmax(kind, m->S));
            m->tc_rnn = m->tc_max_clusters > 0;
        }
        cudaGetLastError();
    }
    if (ok && !m->simt_rnn && !m->tc_rnn) {
        set_err("ffb_model_create: size %d needs the tensor recurrent kernel, which could not be set up", m->S);
        ok = false;
    }
    if (ok && m->simt_rnn && ffb_rnn_prepare(kind, m->S) != 0) {
        set_err("ffb_model_create: recurrent kernel setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        ok = false;
    }
    if (!ok) {
        if (g_err.empty()) set_err("ffb_model_create: device allocation failed");
        ffb_model_destroy(m);
        return nullptr;
    }
    return m;
}

extern "C" int ffb_model_size(const ffb_model *m) { return m ? m->S : 0; }
extern "C" int ffb_model_nparam(const ffb_model *m) { return m ? m->nparam : 0; }
extern "C" int ffb_model_stride(const ffb_model *m) {
    if (!m) return 0;
    int s = 1;
    for (int i = 0; i < m->nconv; i++) s *= m->conv_stride[i];
    return s;
}
extern "C" long ffb_model_nblock(const ffb_model *m, long nsample) {
    if (!m) return -1;
    long t = nsample;
    for (int i = 0; i < m->nconv; i++) {
        if (t < m->conv_winlen[i]) return -1;
        t = (t + m->conv_stride[i] - 1) / m->conv_stride[i];
    }
    return t;
}

// ------------------------------------------------------------------------------------
// The reference's convolution plan (src/layers.c:201-271), replayed with its own integer
// arithmetic for the trailing columns only: which (x_start, tap_lo, ntap) terms does
// column c receive?  Everything before `tail_col0` is the zero-padded window.
struct Term { int x_start, tap_lo, ntap; };

static bool build_conv_tail(int T, int winlen, int stride, ffb::ConvTail *out) {
    const long padL = (winlen - 1) / 2, padR = winlen / 2;
    const long ncol = (T + stride - 1) / stride;
    const long ncolsL = (padL + stride - 1) / stride;          // :229
    const long shiftX = ncolsL * stride - padL;                // :233
    const long nstepC = (winlen + stride - 1) / stride;        // :236
    const long nstepX = stride * nstepC;                       // :237
    const long first = std::max(0L, ncol - FFB_CONV_TAIL);
    std::vector<std::vector<Term>> terms((size_t)(ncol - first));
    auto add = [&](long col, long xs, long tl, long nt) -> bool {
        if (nt <= 0) return true;
        if (col < 0 || col >= ncol || xs < 0 || xs + nt > T) return false;   // the reference itself would leave its matrices
        if (col >= first) terms[(size_t)(col - first)].push_back(Term{(int)xs, (int)tl, (int)nt});
        return true;
    };
    // left edge (:219-226)
    for (long w = 0; w < padL; w += stride)
        if (!add(w / stride, 0, padL - w, winlen - (padL - w))) return false;
    // interior (:239-254): only windows landing in the tail are materialised
    for (long w = 0; w < winlen; w += stride) {
        const long nproc = (T - shiftX - w) / nstepX;          // ifloor on (int)
        const long c_first = ncolsL + w / stride;
        long i0 = 0;
        if (first > c_first) i0 = (first - c_first + nstepC - 1) / nstepC;
        for (long i = i0; i < nproc; i++)
            if (!add(c_first + i * nstepC, shiftX + w + i * nstepX, 0, winlen)) return false;
    }
    // right edge (:257-271)
    const long maxCol = (T - shiftX) / nstepX, rem = (T - shiftX) % nstepX;
    const long colR = ncolsL + nstepC * (maxCol - 1) + rem / stride + 1;
    const long xR = T - winlen + 1;
    const long startR = stride - (padL + T - winlen) % stride - 1;
    for (long w = startR; w < padR; w += stride)
        if (!add(colR + w / stride, xR + w, 0, winlen - (w + 1))) return false;
    // a right-edge column that lands before the window we materialised would be a plan we cannot express
    if (colR + startR / stride < first && startR < padR) return false;

    // first deviating column
    long tail0 = ncol;
    for (long c = first; c < ncol; c++) {
        const auto &tv = terms[(size_t)(c - first)];
        long xs = c * stride - padL, tl = 0, nt = winlen;
        if (xs < 0) { tl = -xs; nt = winlen + xs; xs = 0; }
        if (xs + nt > T) nt = T - xs;
        const bool textbook = tv.size() == 1 && tv[0].x_start == xs && tv[0].tap_lo == tl && tv[0].ntap == nt;
        if (!textbook) { tail0 = c; break; }
    }
    memset(out, 0, sizeof(*out));
    out->tail_col0 = (int)tail0;
    if (tail0 == first && first > 0) {
        // the whole materialised window deviates: cannot prove earlier columns are textbook
        return false;
    }
    for (long c = tail0; c < ncol; c++) {
        const auto &tv = terms[(size_t)(c - first)];
        if (tv.size() > 2) return false;
        for (size_t q = 0; q < tv.size(); q++) {
            out->x_start[c - tail0][q] = tv[q].x_start;
            out->tap_lo[c - tail0][q] = tv[q].tap_lo;
            out->ntap[c - tail0][q] = tv[q].ntap;
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------
struct ffb_ctx {
    ffb_model *m = nullptr;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    // batch plan (host)
    int64_t n_reads = 0, total_samples = 0, total_blocks = 0;
    float temperature = 1.0f;
    uint32_t flags = 0;
    std::vector<int64_t> blk_off;                 // n_reads + 1
    std::vector<int64_t> col_off[FFB_MAX_CONV + 1];   // per conv stage: column offsets (stage 0 = samples)
    std::vector<int32_t> order;
    int max_T[FFB_MAX_CONV + 1] = {0};
    int n_slots = 0;
    bool use_tc_rnn = false;
    int R_tc = 0;                 // reads per cluster of the tensor recurrent kernel = 16 * groups per cluster
    int tc_clusters = 0;          // clusters launched; each of its R_tc/16 slots walks a list of 16-read groups
    std::vector<int32_t> slot_off, slot_list;
    // device
    DevBuf d_sig, d_c[2], d_act[2], d_xin, d_trans, d_tpost, d_fwd, d_tb, d_path, d_qpath, d_score, d_logz, d_trace;
    DevBuf d_geom[FFB_MAX_CONV], d_tails[FFB_MAX_CONV], d_blkoff, d_order, d_keep[FFB_NLAYER];
    DevBuf d_raw, d_rawoff, d_chunkoff, d_mad, d_bounds, d_sigoff;
    DevBuf d_slotoff, d_slotlist;
    DevBuf d_bases, d_quals, d_nbases;   // device-side emission (emit.cu), read n at blk_off[n] + n
    bool want_emit = false;
    bool ff_done = false;         // this forward: the top recurrent layer produced trans
    // pinned staging of the plan arrays (upload_impl) and of the raw phase (offsets out, trim bounds back)
    PinArena h_plan, h_raw;
    cudaEvent_t ev_plan = nullptr;    // recorded behind the last copy out of the arenas: the next batch waits for it first
    bool plan_pending = false;
    cudaEvent_t ev_tail = nullptr;    // recorded when this context's batch enters its last recurrent layer (ffb_model::gate_ev)
    bool gated = false;               // this batch's device work already waits behind the head gate
    // ffb_submit_raw_begin .. ffb_submit_raw_finish
    bool raw_begun = false;
    ffb_raw_batch raw_rb;
    ffb_batch raw_b;
    int64_t *h_bounds = nullptr;      // in h_raw
    DevBuf d_c2hi, d_c2lo;        // tensor-core convolution: fp16 planes of its input in the slot layout (see forward_impl)
    int conv3_fix = 0;            // columns at either end of a read the CUDA-core kernel recomputes
    bool use_tc_conv3 = false;
    DevBuf d_rle;                 // run-length head: shape / scale rows of every block, [Ttot][8]   // device signal preparation (ffb_upload_raw)
    DevBuf d_ahi, d_alo;          // fp16 hi/lo planes of the current layer input (tensor path)
    DevBuf d_ring;                // state-exchange ring of the tensor recurrent kernel (L2-resident)
    // streamed input GEMMs: layer l+1's projection runs on the SMs layer l's recurrence leaves free and consumes
    // its output planes tile by tile (tile order + dependencies per production direction: 0 = forward, 1 = backward)
    DevBuf d_xin2, d_work[2], d_progress;
    bool stream_gemm = false;
    int n_groups = 0;
    float *last_conv = nullptr;   // device pointer of last conv output within d_act/d_c
    cudaEvent_t ev[8] = {nullptr};
    float t_gemm_ms = 0.f, t_rnn_ms = 0.f;
};

extern "C" ffb_ctx *ffb_create(ffb_model *m, void *stream) {
    if (!m) { set_err("ffb_create: NULL model"); return nullptr; }
    CUDA_TRY(cudaSetDevice(m->device), nullptr);
    ffb_ctx *c = new ffb_ctx();
    c->m = m;
    if (stream) {
        c->st = reinterpret_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
            set_err("ffb_create: cudaStreamCreate failed");
            delete c;
            return nullptr;
        }
        c->own_stream = true;
    }
    for (auto &e : c->ev) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_tail, cudaEventDisableTiming);
    return c;
}

extern "C" void ffb_destroy(ffb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->m->device);
    cudaStreamSynchronize(c->st);
    {
        std::lock_guard<std::mutex> lk(c->m->gate_mu);
        if (c->m->gate_owner == c) { c->m->gate_owner = nullptr; c->m->gate_ev = nullptr; }
    }
    if (c->ev_tail) cudaEventDestroy(c->ev_tail);
    DevBuf *all[] = {&c->d_sig, &c->d_c[0], &c->d_c[1], &c->d_act[0], &c->d_act[1], &c->d_xin, &c->d_trans, &c->d_tpost,
                     &c->d_fwd, &c->d_tb, &c->d_path, &c->d_qpath, &c->d_score, &c->d_logz, &c->d_trace, &c->d_blkoff,
                     &c->d_order, &c->d_ahi, &c->d_alo, &c->d_ring, &c->d_xin2, &c->d_work[0], &c->d_work[1], &c->d_progress,
                     &c->d_raw, &c->d_rawoff, &c->d_chunkoff, &c->d_mad, &c->d_bounds, &c->d_sigoff, &c->d_slotoff, &c->d_slotlist, &c->d_rle, &c->d_c2hi, &c->d_c2lo,
                     &c->d_bases, &c->d_quals, &c->d_nbases};
    for (auto *b : all) b->release();
    for (int i = 0; i < FFB_MAX_CONV; i++) { c->d_geom[i].release(); c->d_tails[i].release(); }
    for (int i = 0; i < FFB_NLAYER; i++) c->d_keep[i].release();
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->ev_plan) cudaEventDestroy(c->ev_plan);
    c->h_plan.release(); c->h_raw.release();
    if (c->own_stream) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" int64_t ffb_total_blocks(const ffb_ctx *c) { return c ? c->total_blocks : 0; }
extern "C" int64_t ffb_launch_count(const ffb_ctx *c) { return c ? c->launches : 0; }
extern "C" int ffb_sync(ffb_ctx *c) {
    if (!c) return FFB_ERR_ARG;
    CUDA_TRY(cudaStreamSynchronize(c->st), FFB_ERR_CUDA);
    return FFB_OK;
}

// ---- planning + H2D ------------------------------------------------------------------
#define LAUNCH_RAW(expr)                                                     \
    do {                                                                     \
        const int n_ = (expr);                                               \
        if (n_ < 0) {                                                        \
            set_err("launch failed: %s: %s", #expr, cudaGetErrorString(cudaGetLastError())); \
            return FFB_ERR_CUDA;                                             \
        }                                                                    \
        c->launches += n_;                                                   \
    } while (0)
// ---- schedule of the tensor recurrent kernel (host) ---------------------------------------------------------------
// GROUPS of 16 length-sorted reads; a cluster has G slots, each slot walks a LIST of groups one after the other.
// As few slots per cluster as still fit the batch into one wave of co-resident clusters (the layer is latency-bound:
// more clusters = shorter chains per SM); when the batch has more groups than slots, or the reads differ in length, the
// groups are dealt longest-first to the least-loaded slot (LPT) so every slot runs about the same number of steps.
// blk_off: n_reads + 1 block offsets; sorted: read indices by descending length.
static void plan_groups(const int64_t *blk_off, const int32_t *sorted, int64_t N, int maxc, int gmax, bool can_stream, int csize,
                        int *G_out, int *ncl_out, std::vector<int32_t> &order, std::vector<int32_t> &slot_off,
                        std::vector<int32_t> &slot_list, std::vector<int64_t> &group_start) {
    const int64_t groups = (N + 15) / 16;
    int G = 1, ncl = 0;
    // Which (slots per cluster, clusters)?  The layer lasts makespan(groups dealt longest-first over G * ncl slots) steps and
    // a step costs chain(G): the groups of a CTA share its tensor pipe, MUFU and issue slots, so the step chain grows with
    // the slots in use -- measured 1.72 / 2.0 / 2.26 / 2.7 us per step at 3 / 4 / 5 / 6 slots (profiles/r02_slots_ab.txt), i.e.
    // ~ (3.5 + G) / 8.5 of the five-slot chain, and another 15 % when six slots run on all 15 clusters (the board is power
    // capped: more work in flight lowers the clock).  How much of the streamed input GEMM hides behind the layer did NOT
    // follow the SMs left free in that measurement (six slots x 15 clusters exposed less of it than five x 13), so the
    // model leaves it out.  Candidates are few; each is simulated.  One wave of 16-read groups (BASELINE configs[1]: 64 groups) lands on
    // five slots x 13 clusters as before; 4096 equal reads on six slots x 15 clusters (three rounds instead of four); a
    // ragged batch whose longest read outlasts the average slot on FEWER slots per cluster (its makespan is that read
    // whatever the slot count, so the shorter chain wins: configs[3] runs 24 % faster on three slots than on five).
    {
        std::vector<int64_t> Tg((size_t)groups);
        int64_t blocks = 0;
        for (int64_t g = 0; g < groups; g++) {
            const int32_t rd0 = sorted[g * 16];
            Tg[(size_t)g] = blk_off[rd0 + 1] - blk_off[rd0];
        }
        for (int64_t n = 0; n < N; n++) blocks += blk_off[n + 1] - blk_off[n];
        (void)blocks; (void)can_stream; (void)csize;
        double best = -1.0;
        for (int g = 1; g <= gmax; g++) {
            int tried[4] = {(int)std::min<int64_t>((groups + g - 1) / g, maxc), maxc - 2, maxc - 1, maxc};   // ties: fewer clusters
            for (int k = 0; k < 4; k++) {
                const int c = tried[k];
                if (c < 1) continue;
                bool dup = false;
                for (int q = 0; q < k; q++) dup = dup || tried[q] == c;
                if (dup) continue;
                const int nslot = g * c;
                std::vector<int64_t> load((size_t)nslot, 0);
                for (int64_t i = 0; i < groups; i++) {       // LPT on the (already descending) group lengths
                    size_t b = 0;
                    for (size_t sl = 1; sl < load.size(); sl++)
                        if (load[sl] < load[b]) b = sl;
                    load[b] += Tg[(size_t)i];
                }
                const double makespan = (double)*std::max_element(load.begin(), load.end());
                const double chain = (3.5 + std::max(g, 3)) / 8.5 * ((g >= 6 && g * c > 80) ? 1.15 : 1.0);   // not measured below 3
                const double cost = makespan * chain;
                if (best < 0.0 || cost < best * 0.995) { best = cost; G = g; ncl = c; }      // ties: fewer slots, fewer clusters
            }
        }
        if (groups == 0) { G = 1; ncl = 0; }
    }
    if (getenv("FFB_TC_SLOTS")) {             // experiments: slots per cluster
        G = std::max(1, std::min(atoi(getenv("FFB_TC_SLOTS")), gmax));
        ncl = (int)std::min<int64_t>((groups + G - 1) / G, maxc);
    }
    if (getenv("FFB_TC_CLUSTERS")) ncl = std::max(1, std::min(atoi(getenv("FFB_TC_CLUSTERS")), maxc));
    if (groups == 0) ncl = 0;
    order.assign((size_t)(groups * 16), -1);
    std::copy(sorted, sorted + N, order.begin());
    const int nslot = ncl * G;
    std::vector<std::vector<int32_t>> lists((size_t)nslot);
    std::vector<int64_t> load((size_t)nslot, 0);
    group_start.assign((size_t)groups, 0);
    for (int64_t g = 0; g < groups; g++) {          // groups are already in descending order of their longest read
        const int32_t rd0 = order[(size_t)g * 16];
        const int64_t Tg = blk_off[rd0 + 1] - blk_off[rd0];
        int best = 0;
        if (getenv("FFB_TC_ROUND_ROBIN")) {         // A/B: groups dealt in sorted order, as successive waves would run them
            best = (int)(g % nslot);
        } else {
            for (int sl = 1; sl < nslot; sl++)
                if (load[(size_t)sl] < load[(size_t)best]) best = sl;
        }
        group_start[(size_t)g] = load[(size_t)best];
        lists[(size_t)best].push_back((int32_t)g);
        load[(size_t)best] += Tg;
    }
    slot_off.assign((size_t)nslot + 1, 0);
    slot_list.clear();
    for (int sl = 0; sl < nslot; sl++) {
        slot_off[(size_t)sl + 1] = slot_off[(size_t)sl] + (int32_t)lists[(size_t)sl].size();
        slot_list.insert(slot_list.end(), lists[(size_t)sl].begin(), lists[(size_t)sl].end());
    }
    *G_out = G; *ncl_out = ncl;
}

// The same planner for callers / tests without a device: T[n] = blocks of read n.  Outputs: order (16 * groups entries,
// -1 padded), slot_off (n_clusters * G + 1), slot_list (groups).  Returns the number of groups.
extern "C" int64_t ffb_plan_schedule(const int64_t *T, int64_t n_reads, int max_clusters, int slots_max, int can_stream,
                                     int32_t *order, int32_t *slot_off, int32_t *slot_list, int *n_clusters, int *slots) {
    if (!T || n_reads < 0 || max_clusters < 1 || slots_max < 1 || !n_clusters || !slots) return -1;
    std::vector<int64_t> off((size_t)n_reads + 1, 0);
    for (int64_t n = 0; n < n_reads; n++) off[(size_t)n + 1] = off[(size_t)n] + std::max<int64_t>(T[n], 0);
    std::vector<int32_t> idx((size_t)n_reads);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return (off[a + 1] - off[a]) > (off[b + 1] - off[b]); });
    std::vector<int32_t> o, so, sl;
    std::vector<int64_t> gs;
    plan_groups(off.data(), idx.data(), n_reads, max_clusters, slots_max, can_stream != 0, 8, slots, n_clusters, o, so, sl, gs);
    if (order) std::copy(o.begin(), o.end(), order);
    if (slot_off) std::copy(so.begin(), so.end(), slot_off);
    if (slot_list) std::copy(sl.begin(), sl.end(), slot_list);
    return (int64_t)sl.size();
}

// copy_signal = false: the normalised signal is produced on the device (ffb_upload_raw), only the plan is made here
static int upload_impl(ffb_ctx *c, const ffb_batch *b, bool copy_signal) {
    if (!c || !b || (copy_signal && !b->signal) || !b->sig_off || b->n_reads < 0) { set_err("ffb_upload: bad arguments"); return FFB_ERR_ARG; }
    ffb_model *m = c->m;
    CUDA_TRY(cudaSetDevice(m->device), FFB_ERR_CUDA);
    const int64_t N = b->n_reads;
    if (N > 0x7fffffff) { set_err("ffb_upload: too many reads"); return FFB_ERR_ARG; }
    for (int64_t n = 0; n < N; n++)
        if (b->sig_off[n + 1] < b->sig_off[n]) { set_err("ffb_upload: sig_off must be non-decreasing (read %lld)", (long long)n); return FFB_ERR_ARG; }
    if (b->sig_off[N] - b->sig_off[0] > ((int64_t)1 << 40)) { set_err("ffb_upload: batch too large"); return FFB_ERR_ARG; }
    c->n_reads = N; c->temperature = b->temperature; c->flags = b->flags;
    c->total_samples = b->sig_off[N] - b->sig_off[0];

    // per-stage column counts: stage 0 = samples, stage i+1 = output of conv i
    for (int s = 0; s <= m->nconv; s++) { c->col_off[s].assign(N + 1, 0); c->max_T[s] = 0; }
    std::vector<ffb::ReadGeom> geom[FFB_MAX_CONV];
    std::vector<ffb::ConvTail> tails[FFB_MAX_CONV];
    std::map<int, int> tail_id[FFB_MAX_CONV];
    for (int i = 0; i < m->nconv; i++) geom[i].resize((size_t)N);
    for (int64_t n = 0; n < N; n++) {
        long T = (long)(b->sig_off[n + 1] - b->sig_off[n]);
        const bool ok = ffb_model_nblock(m, T) > 0;
        if (!ok) T = 0;   // rejected read: zero columns everywhere (reference would underflow, layers.c:262)
        c->col_off[0][n + 1] = c->col_off[0][n] + (b->sig_off[n + 1] - b->sig_off[n]);   // samples are uploaded as they are
        c->max_T[0] = std::max<int>(c->max_T[0], (int)T);
        for (int i = 0; i < m->nconv; i++) {
            const long To = ok ? (T + m->conv_stride[i] - 1) / m->conv_stride[i] : 0;
            ffb::ReadGeom g;
            g.in_off = c->col_off[i][n]; g.out_off = c->col_off[i + 1][n];
            g.T_in = (int)T; g.T_out = (int)To; g.tail_id = 0; g.pad = 0; g.plane_off = g.out_off;
            if (ok) {
                auto it = tail_id[i].find((int)T);
                if (it == tail_id[i].end()) {
                    ffb::ConvTail tl;
                    bool have = false;
                    {
                        std::lock_guard<std::mutex> lk(m->mu);
                        auto ci = m->tail_cache.find({i, (int)T});
                        if (ci != m->tail_cache.end()) { tl = ci->second; have = true; }
                    }
                    if (!have) {
                        if (!build_conv_tail((int)T, m->conv_winlen[i], m->conv_stride[i], &tl)) {
                            set_err("ffb_upload: convolution %d edge plan not expressible for T=%ld", i, T);
                            return FFB_ERR_UNSUPPORTED;
                        }
                        std::lock_guard<std::mutex> lk(m->mu);
                        if (m->tail_cache.size() >= FFB_TAIL_CACHE_MAX) m->tail_cache.clear();   // bounded: a plan is cheap to rebuild
                        m->tail_cache[{i, (int)T}] = tl;
                    }
                    const int id = (int)tails[i].size();
                    tails[i].push_back(tl);
                    it = tail_id[i].emplace((int)T, id).first;
                }
                g.tail_id = it->second;
            }
            geom[i][(size_t)n] = g;
            c->col_off[i + 1][n + 1] = c->col_off[i + 1][n] + To;
            c->max_T[i + 1] = std::max<int>(c->max_T[i + 1], (int)To);
            T = To;
        }
    }
    c->blk_off = c->col_off[m->nconv];
    c->total_blocks = c->blk_off[N];
    // kernels index blocks (and blocks * nparam / 16) with 32 bits
    if (c->total_blocks > 0x7fffffff || c->col_off[0][N] > ((int64_t)1 << 40)) {
        set_err("ffb_upload: %lld blocks in one batch exceed the 2^31 - 1 the kernels index", (long long)c->total_blocks);
        return FFB_ERR_ARG;
    }
    if (b->blk_off) memcpy(b->blk_off, c->blk_off.data(), sizeof(int64_t) * (size_t)(N + 1));
    // tensor-core last convolution: its input planes use a SLOT layout -- read n's columns start at column
    // stride * blk_off[n], so that output block r (of the whole batch) is the window starting stride * r - padL columns
    // into the planes and ONE overlapping-row tensor map describes the im2col matrix of the batch
    c->use_tc_conv3 = m->tc_conv3 && m->tc_gemm && !(c->flags & (FFB_FLAG_FP32_SIMT | FFB_FLAG_FP32_CONV)) && c->total_blocks > 0;
    c->conv3_fix = 0;
    if (c->use_tc_conv3) {
        const int last = m->nconv - 1;
        int fix = 2;                                   // windows of the first / last two columns reach into a neighbouring read
        for (int64_t n = 0; n < N; n++) {
            geom[last - 1][(size_t)n].plane_off = (int64_t)m->conv_stride[last] * c->blk_off[n];
            const ffb::ReadGeom &g = geom[last][(size_t)n];
            if (g.T_out > 0) fix = std::max(fix, g.T_out - tails[last][(size_t)g.tail_id].tail_col0);   // the reference's edge plan
        }
        c->conv3_fix = 16 * ((fix + 15) / 16);
    }

    // length-sorted slots for the recurrent kernel (descending, stable)
    int R = m->simt_rnn ? ffb_rnn_reads_per_cluster(m->kind, m->S) : 16;
    c->use_tc_rnn = m->tc_rnn && !(c->flags & FFB_FLAG_FP32_SIMT) && getenv("FFB_NO_TC_RNN") == nullptr;
    if (!c->use_tc_rnn && !m->simt_rnn) {
        set_err("ffb_upload: size %d has no fp32 CUDA-core recurrence (FFB_FLAG_FP32_SIMT / FFB_NO_TC_RNN not available)", m->S);
        return FFB_ERR_UNSUPPORTED;
    }
    std::vector<int32_t> idx((size_t)N);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t bb) {
        return (c->blk_off[a + 1] - c->blk_off[a]) > (c->blk_off[bb + 1] - c->blk_off[bb]);
    });
    std::vector<int64_t> group_start;   // steps a group's slot has already run when the group begins
    c->slot_off.clear(); c->slot_list.clear(); c->tc_clusters = 0;
    if (c->use_tc_rnn) {
        const bool can_stream = ffb_gemm_tc_stream_supported(m->G * m->S, m->S) != 0;
        int G = 1, ncl = 0;
        plan_groups(c->blk_off.data(), idx.data(), N, std::max(m->tc_max_clusters, 1), ffb_rnn_tc_rmax(m->kind, m->S) / 16, can_stream,
                    ffb_rnn_tc_cluster_size(m->kind, m->S), &G, &ncl, c->order, c->slot_off, c->slot_list, group_start);
        c->R_tc = G * 16;
        c->tc_clusters = ncl;
        c->n_slots = (int)c->order.size();
    } else {
        c->n_slots = (int)(((N + R - 1) / R) * R);
        c->order.assign((size_t)c->n_slots, -1);
        std::copy(idx.begin(), idx.end(), c->order.begin());
    }

    // ---- streamed GEMM plan: when may each tile of the next layer's input be loaded? ----
    const int64_t Tt = c->total_blocks, S = m->S, G = m->G, nr = m->nparam;
    std::vector<GemmWork> work[2];
    c->stream_gemm = c->use_tc_rnn && !(c->flags & FFB_FLAG_KEEP_LAYERS) && getenv("FFB_NO_STREAM_GEMM") == nullptr &&
                     ffb_gemm_tc_stream_supported((int)(G * S), (int)S) && Tt > 0;
    if (c->stream_gemm) {
        const int R16 = 16, P = FFB_RNN_PUBLISH_PERIOD;
        c->n_groups = c->n_slots / R16;
        std::vector<int32_t> group_of((size_t)N, -1), group_T((size_t)c->n_groups, 0);
        for (int sl = 0; sl < c->n_slots; sl++) {
            const int32_t rd = c->order[(size_t)sl];
            if (rd < 0) continue;
            group_of[(size_t)rd] = sl / R16;
            if (sl % R16 == 0) group_T[(size_t)(sl / R16)] = (int32_t)(c->blk_off[rd + 1] - c->blk_off[rd]);
        }
        const int64_t TR = ffb_gemm_tc_stream_tile_rows((int)S);
        const int64_t n_tiles = (Tt + TR - 1) / TR;
        const int arrivals = 4 * ffb_rnn_tc_cluster_size(m->kind, m->S);   // gate warps per group and cluster: C CTAs x 4 quadrants
        for (int dir = 0; dir < 2; dir++) {
            work[dir].resize((size_t)n_tiles);
            std::vector<int32_t> ready((size_t)n_tiles, 0);
            int64_t rd = 0;
            for (int64_t k = 0; k < n_tiles; k++) {
                const int64_t r0 = k * TR, r1 = std::min<int64_t>(r0 + TR, Tt) - 1;
                GemmWork w; w.tile = (int32_t)k; w.pad = 0;
                for (int d = 0; d < 3; d++) { w.idx[d] = -1; w.cnt[d] = 0; }
                while (c->blk_off[rd + 1] <= r0) rd++;          // first read with a row in the tile
                int nd = 0, worst = 0; bool overflow = false;
                for (int64_t n = rd; n < N && c->blk_off[n] <= r1; n++) {
                    const int64_t T = c->blk_off[n + 1] - c->blk_off[n];
                    if (T == 0) continue;
                    const int64_t ta = std::max<int64_t>(r0, c->blk_off[n]) - c->blk_off[n];
                    const int64_t tb = std::min<int64_t>(r1, c->blk_off[n + 1] - 1) - c->blk_off[n];
                    const int64_t need = dir == 0 ? tb + 1 : T - ta;      // steps of the producing layer
                    const int32_t g = group_of[(size_t)n];
                    const int events_total = (group_T[(size_t)g] + P - 1) / P;
                    const int ev = (int)std::min<int64_t>((need + P - 1) / P, events_total);
                    // step (of its slot) at which the producing layer publishes this event
                    worst = (int)std::min<int64_t>(std::max<int64_t>(worst, group_start[(size_t)g] + (int64_t)ev * P), 0x7ffffffe);
                    if (nd < 3) { w.idx[nd] = g; w.cnt[nd] = ev * arrivals; nd++; }
                    else overflow = true;
                }
                if (overflow) {
                    // more than three reads in one tile (very short reads): wait for the whole layer instead --
                    // the counter at index n_groups counts CTAs that have finished it
                    for (int d = 0; d < 3; d++) { w.idx[d] = -1; w.cnt[d] = 0; }
                    w.idx[0] = c->n_groups; w.cnt[0] = c->tc_clusters * ffb_rnn_tc_cluster_size(m->kind, m->S);
                    worst = 0x7fffffff;
                }
                work[dir][(size_t)k] = w;
                ready[(size_t)k] = worst;
            }
            std::stable_sort(work[dir].begin(), work[dir].end(),
                             [&](const GemmWork &a, const GemmWork &b) { return ready[(size_t)a.tile] < ready[(size_t)b.tile]; });
        }
    }

    // ---- device workspaces ----
    bool ok = true;
    ok &= c->d_sig.reserve(sizeof(float) * (size_t)std::max<int64_t>(c->total_samples, 1)) == 0;
    for (int i = 0; i + 1 < m->nconv; i++)
        ok &= c->d_c[i].reserve(sizeof(float) * (size_t)std::max<int64_t>(c->col_off[i + 1][N] * m->conv_nfilter[i], 1)) == 0;
    // fp32 activations [Tt][S] exist only off the tensor path (there the layers hand each other fp16 hi/lo planes) or when
    // the caller wants to look at them
    const bool tensor_path = m->tc_gemm && !(c->flags & FFB_FLAG_FP32_SIMT) && c->use_tc_rnn && m->tc_ff && getenv("FFB_NO_TC_FF") == nullptr;
    const size_t act_bytes = (tensor_path && !(c->flags & FFB_FLAG_KEEP_LAYERS)) ? 256 : sizeof(float) * (size_t)std::max<int64_t>(Tt * S, 1);
    ok &= c->d_act[0].reserve(act_bytes) == 0;
    ok &= c->d_act[1].reserve(act_bytes) == 0;
    ok &= c->d_xin.reserve(sizeof(float) * (size_t)std::max<int64_t>(Tt * G * S, 1)) == 0;
    if (m->tc_gemm && !(c->flags & FFB_FLAG_FP32_SIMT)) {
        ok &= c->d_ahi.reserve(2 * (size_t)std::max<int64_t>(Tt * S, 1)) == 0;
        ok &= c->d_alo.reserve(2 * (size_t)std::max<int64_t>(Tt * S, 1)) == 0;
    }
    if (c->stream_gemm) {
        for (int dir = 0; dir < 2; dir++) ok &= c->d_work[dir].reserve(sizeof(GemmWork) * work[dir].size()) == 0;
        ok &= c->d_progress.reserve(sizeof(int) * (size_t)FFB_NLAYER * (c->n_groups + 1 + 16)) == 0;
    }
    if (c->use_tc_rnn)
    {
        ok &= c->d_ring.reserve(std::max<size_t>(ffb_rnn_tc_ring_bytes(m->kind, m->S, c->tc_clusters, c->R_tc), 16)) == 0;
        ok &= c->d_slotoff.reserve(sizeof(int32_t) * std::max<size_t>(c->slot_off.size(), 1)) == 0;
        ok &= c->d_slotlist.reserve(sizeof(int32_t) * std::max<size_t>(c->slot_list.size(), 1)) == 0;
    }
    ok &= c->d_trans.reserve(sizeof(float) * (size_t)std::max<int64_t>(Tt * nr, 1)) == 0;
    if (!(c->flags & FFB_FLAG_VITERBI_ONLY)) {
        ok &= c->d_tpost.reserve(sizeof(float) * (size_t)std::max<int64_t>(Tt * nr, 1)) == 0;
        ok &= c->d_fwd.reserve(2 * sizeof(float) * (size_t)std::max<int64_t>((Tt + N) * m->nstate, 1)) == 0;   // forward + backward vectors
    }
    ok &= c->d_tb.reserve(sizeof(uint64_t) * (size_t)std::max<int64_t>(Tt, 1)) == 0;
    ok &= c->d_path.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(Tt + N, 1)) == 0;
    ok &= c->d_qpath.reserve(sizeof(float) * (size_t)std::max<int64_t>(Tt + N, 1)) == 0;
    ok &= c->d_score.reserve(sizeof(float) * (size_t)std::max<int64_t>(N, 1)) == 0;
    c->want_emit = b->bases && b->quals && b->nbases && !m->head;
    if (c->want_emit) {
        ok &= c->d_bases.reserve((size_t)std::max<int64_t>(Tt + N, 1)) == 0;
        ok &= c->d_quals.reserve((size_t)std::max<int64_t>(Tt + N, 1)) == 0;
        ok &= c->d_nbases.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(N, 1)) == 0;
    }
    ok &= c->d_logz.reserve(sizeof(double) * (size_t)std::max<int64_t>(N, 1)) == 0;
    if (c->flags & FFB_FLAG_WANT_TRACE) ok &= c->d_trace.reserve((size_t)std::max<int64_t>((Tt + N) * m->nstate, 1)) == 0;
    if (c->flags & FFB_FLAG_KEEP_LAYERS)
        for (int l = 0; l < FFB_NLAYER; l++) ok &= c->d_keep[l].reserve(sizeof(float) * (size_t)std::max<int64_t>(Tt * S, 1)) == 0;
    if (c->use_tc_conv3) {
        const int last = m->nconv - 1;
        const size_t halfs = (size_t)((m->conv_winlen[last] - 1) / 2) * m->conv_nf[last] +
                             (size_t)m->conv_stride[last] * (size_t)Tt * m->conv_nf[last] + (size_t)m->c3_kp + 64;
        ok &= c->d_c2hi.reserve(2 * halfs) == 0;
        ok &= c->d_c2lo.reserve(2 * halfs) == 0;
    }
    if (m->head) ok &= c->d_rle.reserve(sizeof(float) * 8 * (size_t)std::max<int64_t>(Tt, 1)) == 0;
    ok &= c->d_blkoff.reserve(sizeof(int64_t) * (size_t)(N + 1)) == 0;
    ok &= c->d_order.reserve(sizeof(int32_t) * (size_t)std::max(c->n_slots, 1)) == 0;
    for (int i = 0; i < m->nconv; i++) {
        ok &= c->d_geom[i].reserve(sizeof(ffb::ReadGeom) * (size_t)std::max<int64_t>(N, 1)) == 0;
        ok &= c->d_tails[i].reserve(sizeof(ffb::ConvTail) * std::max<size_t>(tails[i].size(), 1)) == 0;
    }
    if (!ok) { set_err("ffb_upload: out of device memory for %lld reads / %lld blocks", (long long)N, (long long)Tt); return FFB_ERR_NOMEM; }

    // ---- H2D (signal is the only bulk input: 4 bytes per raw sample) ----
    // the plan arrays go through the pinned arena: no copy waits for the stream, nothing to synchronise afterwards
    if (c->plan_pending) { CUDA_TRY(cudaEventSynchronize(c->ev_plan), FFB_ERR_CUDA); c->plan_pending = false; }   // an earlier batch still reading it
    {
        size_t need = 256 + sizeof(int64_t) * (size_t)(N + 1) + sizeof(int32_t) * ((size_t)c->n_slots + c->slot_off.size() + c->slot_list.size());
        for (int i = 0; i < m->nconv; i++) need += 64 + sizeof(ffb::ReadGeom) * (size_t)N + sizeof(ffb::ConvTail) * tails[i].size();
        for (int dir = 0; dir < 2; dir++) need += 64 + sizeof(GemmWork) * work[dir].size();
        if (c->h_plan.reserve(need + 1024) != 0) { set_err("ffb_upload: out of pinned host memory"); return FFB_ERR_NOMEM; }
    }
    auto h2d = [&](void *dst, const void *src, size_t bytes) -> bool {
        if (bytes == 0) return true;
        void *stage = c->h_plan.put(src, bytes);
        return stage && cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyHostToDevice, c->st) == cudaSuccess;
    };
    if (c->total_samples > 0 && copy_signal)
        CUDA_TRY(cudaMemcpyAsync(c->d_sig.p, b->signal + b->sig_off[0], sizeof(float) * (size_t)c->total_samples,
                                 cudaMemcpyHostToDevice, c->st), FFB_ERR_CUDA);
    bool up = h2d(c->d_blkoff.p, c->blk_off.data(), sizeof(int64_t) * (size_t)(N + 1));
    if (c->n_slots > 0) up = up && h2d(c->d_order.p, c->order.data(), sizeof(int32_t) * (size_t)c->n_slots);
    if (!c->slot_off.empty()) {
        up = up && h2d(c->d_slotoff.p, c->slot_off.data(), sizeof(int32_t) * c->slot_off.size());
        up = up && h2d(c->d_slotlist.p, c->slot_list.data(), sizeof(int32_t) * c->slot_list.size());
    }
    for (int i = 0; i < m->nconv; i++) {
        if (N > 0) up = up && h2d(c->d_geom[i].p, geom[i].data(), sizeof(ffb::ReadGeom) * (size_t)N);
        up = up && h2d(c->d_tails[i].p, tails[i].data(), sizeof(ffb::ConvTail) * tails[i].size());
    }
    if (c->stream_gemm)
        for (int dir = 0; dir < 2; dir++) up = up && h2d(c->d_work[dir].p, work[dir].data(), sizeof(GemmWork) * work[dir].size());
    if (!up) { set_err("ffb_upload: plan upload failed: %s", cudaGetErrorString(cudaGetLastError())); return FFB_ERR_CUDA; }
    CUDA_TRY(cudaEventRecord(c->ev_plan, c->st), FFB_ERR_CUDA);
    c->plan_pending = true;
    return FFB_OK;
}

extern "C" int ffb_upload(ffb_ctx *c, const ffb_batch *b) { return upload_impl(c, b, true); }

// Size the context's grow-only workspaces for batches of n_reads reads of samples_per_read samples before the first real
// batch arrives (cudaMalloc of gigabytes is slow and serialises across the threads of a process): plans a dummy uniform
// batch, which reserves everything, and uploads nothing but its plan.
extern "C" int ffb_reserve(ffb_ctx *c, int64_t n_reads, int64_t samples_per_read, uint32_t flags) {
    if (!c || n_reads < 0 || samples_per_read < 0) return FFB_ERR_ARG;
    std::vector<int64_t> off((size_t)n_reads + 1);
    for (int64_t n = 0; n <= n_reads; n++) off[(size_t)n] = n * samples_per_read;
    ffb_batch plan;
    memset(&plan, 0, sizeof plan);
    plan.sig_off = off.data(); plan.n_reads = n_reads; plan.temperature = 1.0f; plan.flags = flags;
    char dummy = 0;
    plan.bases = &dummy; plan.quals = &dummy; int32_t nb = 0; plan.nbases = &nb;      // the emission buffers too
    int r = upload_impl(c, &plan, false);
    if (r != FFB_OK) return r;
    const int64_t total = n_reads * samples_per_read;
    const bool ok = c->d_raw.reserve(sizeof(float) * (size_t)std::max<int64_t>(total, 1)) == 0 &&
                    c->d_mad.reserve(sizeof(float) * (size_t)std::max<int64_t>(total / 2 + n_reads, 1)) == 0;
    CUDA_TRY(cudaStreamSynchronize(c->st), FFB_ERR_CUDA);
    c->n_reads = 0; c->total_blocks = 0; c->total_samples = 0; c->want_emit = false;     // nothing to run or download
    return ok ? FFB_OK : FFB_ERR_NOMEM;
}

// ---- raw reads: trimming + normalisation on the device, then the same plan -----------------------
// (reference src/flappie.c:251-259; kernels in signal.cu)
// Two halves, because the plan needs the trimmed lengths back from the device: `begin` enqueues the raw upload, the chunk
// MADs, the trim bounds and their copy back, and returns; `finish` waits for the bounds, plans and enqueues the rest.  A host
// thread can do something useful in between (the command line reads the next window's files).
// Head gate, see ffb_model.  A scheduling hint only: results do not depend on it (FFB_NO_HEAD_GATE=1 turns it off).
static bool head_gate_enabled() {
    static const bool on = getenv("FFB_NO_HEAD_GATE") == nullptr;
    return on;
}
// FFB_HEAD_GATE=conv: only the network waits (the trimming kernels run at once, so the host has the trimmed lengths and
// the plan early); default: the trimming kernels wait too
static bool head_gate_prep() {
    static const bool on = !(getenv("FFB_HEAD_GATE") && strcmp(getenv("FFB_HEAD_GATE"), "conv") == 0);
    return on;
}
static void head_gate_wait(ffb_ctx *c) {
    if (c->gated || !head_gate_enabled()) return;
    c->gated = true;
    std::lock_guard<std::mutex> lk(c->m->gate_mu);      // held across the call: the owner cannot destroy the event under it
    if (c->m->gate_ev && c->m->gate_owner != c) cudaStreamWaitEvent(c->st, c->m->gate_ev, 0);
}
static void head_gate_publish(ffb_ctx *c) {
    if (!head_gate_enabled() || !c->ev_tail) return;
    if (cudaEventRecord(c->ev_tail, c->st) != cudaSuccess) { cudaGetLastError(); return; }
    std::lock_guard<std::mutex> lk(c->m->gate_mu);
    c->m->gate_ev = c->ev_tail;
    c->m->gate_owner = c;
}

static int upload_raw_begin(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b) {
    if (!c || !rb || !b || !rb->raw || !rb->raw_off || rb->n_reads < 0 || rb->n_reads != b->n_reads) {
        set_err("ffb_upload_raw: bad arguments");
        return FFB_ERR_ARG;
    }
    if (rb->varseg_chunk < 2 || rb->varseg_chunk > FFB_MAX_VARSEG_CHUNK || rb->varseg_thresh < 0.0f || rb->varseg_thresh > 1.0f ||
        rb->trim_start < 0 || rb->trim_end < 0) {
        set_err("ffb_upload_raw: varseg_chunk must be in [2, %d], varseg_thresh in [0, 1], trims >= 0", FFB_MAX_VARSEG_CHUNK);
        return FFB_ERR_UNSUPPORTED;
    }
    ffb_model *m = c->m;
    CUDA_TRY(cudaSetDevice(m->device), FFB_ERR_CUDA);
    const int64_t N = rb->n_reads;
    if (N > 0x7fffffff) return FFB_ERR_ARG;
    const int chunk = (int)rb->varseg_chunk;
    const int64_t total_raw = rb->raw_off[N] - rb->raw_off[0];
    std::vector<int64_t> roff((size_t)N + 1), coff((size_t)N + 1, 0);
    for (int64_t n = 0; n <= N; n++) roff[(size_t)n] = rb->raw_off[n] - rb->raw_off[0];
    for (int64_t n = 0; n < N; n++) {
        if (roff[(size_t)n + 1] < roff[(size_t)n]) { set_err("ffb_upload_raw: raw_off must be non-decreasing"); return FFB_ERR_ARG; }
        coff[(size_t)n + 1] = coff[(size_t)n] + (roff[(size_t)n + 1] - roff[(size_t)n]) / chunk;
    }
    const int64_t total_chunks = coff[(size_t)N];
    bool ok = c->d_raw.reserve(sizeof(float) * (size_t)std::max<int64_t>(total_raw, 1)) == 0 &&
              c->d_rawoff.reserve(sizeof(int64_t) * (size_t)(N + 1)) == 0 && c->d_chunkoff.reserve(sizeof(int64_t) * (size_t)(N + 1)) == 0 &&
              c->d_mad.reserve(sizeof(float) * (size_t)std::max<int64_t>(total_chunks, 1)) == 0 &&
              c->d_bounds.reserve(sizeof(int64_t) * 2 * (size_t)std::max<int64_t>(N, 1)) == 0 &&
              c->d_sigoff.reserve(sizeof(int64_t) * (size_t)(N + 1)) == 0;
    if (!ok) { set_err("ffb_upload_raw: out of device memory"); return FFB_ERR_NOMEM; }
    cudaStream_t st = c->st;
    if (c->plan_pending) { CUDA_TRY(cudaEventSynchronize(c->ev_plan), FFB_ERR_CUDA); c->plan_pending = false; }
    if (c->h_raw.reserve(sizeof(int64_t) * (5 * (size_t)N + 8) + 256) != 0) { set_err("ffb_upload_raw: out of pinned host memory"); return FFB_ERR_NOMEM; }
    const int64_t *h_roff = (const int64_t *)c->h_raw.put(roff.data(), sizeof(int64_t) * (size_t)(N + 1));
    const int64_t *h_coff = (const int64_t *)c->h_raw.put(coff.data(), sizeof(int64_t) * (size_t)(N + 1));
    c->h_bounds = (int64_t *)c->h_raw.take(sizeof(int64_t) * 2 * (size_t)std::max<int64_t>(N, 1));
    if (!h_roff || !h_coff || !c->h_bounds) return FFB_ERR_NOMEM;
    if (N > 0) {
        if (total_raw > 0)
            CUDA_TRY(cudaMemcpyAsync(c->d_raw.p, rb->raw + rb->raw_off[0], sizeof(float) * (size_t)total_raw, cudaMemcpyHostToDevice, st), FFB_ERR_CUDA);
        CUDA_TRY(cudaMemcpyAsync(c->d_rawoff.p, h_roff, sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, st), FFB_ERR_CUDA);
        CUDA_TRY(cudaMemcpyAsync(c->d_chunkoff.p, h_coff, sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, st), FFB_ERR_CUDA);
        if (head_gate_prep()) head_gate_wait(c);   // the copies may run ahead; the kernels wait for the batch in flight to reach its last layer
        LAUNCH_RAW(ffb_launch_chunk_mad(c->d_raw.as<float>(), c->d_rawoff.as<int64_t>(), c->d_chunkoff.as<int64_t>(), (int)N, chunk,
                                        total_chunks, c->d_mad.as<float>(), st));
        LAUNCH_RAW(ffb_launch_trim_bounds(c->d_mad.as<float>(), c->d_rawoff.as<int64_t>(), c->d_chunkoff.as<int64_t>(), (int)N, chunk,
                                          rb->varseg_thresh, rb->trim_start, rb->trim_end, c->d_bounds.as<int64_t>(), st));
        CUDA_TRY(cudaMemcpyAsync(c->h_bounds, c->d_bounds.p, sizeof(int64_t) * 2 * (size_t)N, cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
    }
    c->raw_rb = *rb; c->raw_b = *b;
    c->raw_begun = true;
    return FFB_OK;
}

static int upload_raw_finish(ffb_ctx *c) {
    if (!c || !c->raw_begun) { set_err("ffb_submit_raw_finish: no batch was begun on this context"); return FFB_ERR_ARG; }
    c->raw_begun = false;
    ffb_model *m = c->m;
    CUDA_TRY(cudaSetDevice(m->device), FFB_ERR_CUDA);
    const ffb_raw_batch *rb = &c->raw_rb;
    const int64_t N = rb->n_reads;
    cudaStream_t st = c->st;
    if (N > 0) CUDA_TRY(cudaStreamSynchronize(st), FFB_ERR_CUDA);     // the plan below needs the trimmed lengths
    std::vector<int64_t> soff((size_t)N + 1, 0);
    for (int64_t n = 0; n < N; n++) {
        const int64_t s = c->h_bounds[2 * (size_t)n], e = c->h_bounds[2 * (size_t)n + 1];
        soff[(size_t)n + 1] = soff[(size_t)n] + (s < e ? e - s : 0);    // start >= end: the reference drops the read (flappie_common.c:22-25)
        if (rb->start) rb->start[n] = s;
        if (rb->end) rb->end[n] = e;
    }
    ffb_batch plan = c->raw_b;
    plan.signal = nullptr;
    plan.sig_off = soff.data();
    if (rb->delta != 0.0f) plan.flags |= FFB_FLAG_FP32_CONV;      // unnormalised delta samples: see flappie_b200.h
    const int r = upload_impl(c, &plan, false);
    if (r != FFB_OK) return r;
    if (N > 0) {
        // soff rides in what is left of the raw arena (its first copies were consumed before the synchronisation above)
        const int64_t *h_soff = (const int64_t *)c->h_raw.put(soff.data(), sizeof(int64_t) * (size_t)(N + 1));
        if (!h_soff) return FFB_ERR_NOMEM;
        CUDA_TRY(cudaMemcpyAsync(c->d_sigoff.p, h_soff, sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, st), FFB_ERR_CUDA);
        LAUNCH_RAW(ffb_launch_normalise(c->d_raw.as<float>(), c->d_rawoff.as<int64_t>(), c->d_bounds.as<int64_t>(), c->d_sigoff.as<int64_t>(),
                                        (int)N, rb->delta, c->d_sig.as<float>(), st));
        CUDA_TRY(cudaEventRecord(c->ev_plan, st), FFB_ERR_CUDA);     // behind the last copy out of either arena
        c->plan_pending = true;
    }
    return FFB_OK;
}

extern "C" int ffb_upload_raw(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b) {
    const int r = upload_raw_begin(c, rb, b);
    return r != FFB_OK ? r : upload_raw_finish(c);
}

// ---- all kernels of the path ------------------------------------------------------------
#define LAUNCH(expr)                                                         \
    do {                                                                     \
        const int n_ = (expr);                                               \
        if (n_ < 0) {                                                        \
            set_err("launch failed: %s: %s", #expr, cudaGetErrorString(cudaGetLastError())); \
            return FFB_ERR_CUDA;                                             \
        }                                                                    \
        c->launches += n_;                                                   \
    } while (0)

#ifdef FFB_RNN_PROFILE
static constexpr bool kProfileBuild = true;      // group events are recorded in the ordinary (streamed) schedule too: tools/step_timeline.py
#else
static constexpr bool kProfileBuild = false;
#endif
static int forward_impl(ffb_ctx *c, bool timed) {
    ffb_model *m = c->m;
    const int64_t N = c->n_reads, Tt = c->total_blocks;
    const int S = m->S, G = m->G, nr = m->nparam;
    cudaStream_t st = c->st;
    if (N == 0 || Tt == 0) return FFB_OK;
    if (N > 0x7fffffff || Tt > 0x7fffffff) return FFB_ERR_ARG;   // kernels index blocks with 32 bits
    if (!timed) head_gate_wait(c);      // no-op when the raw upload of this batch already waited
    c->gated = false;
    if (timed || kProfileBuild) cudaEventRecord(c->ev[0], st);
    // ---- convolutions (features_from_raw folded into the first load) ----
    const float *cur = c->d_sig.as<float>();
    const bool keep = (c->flags & FFB_FLAG_KEEP_LAYERS) != 0;
    const bool tc_gemm = m->tc_gemm && !(c->flags & FFB_FLAG_FP32_SIMT);
    for (int i = 0; i < m->nconv; i++) {
        const bool lastc = (i + 1 == m->nconv);
        float *out = lastc ? c->d_act[0].as<float>() : c->d_c[i].as<float>();
        const int act = (m->kind == FFB_KIND_GRU) ? FFB_ACT_TANH : FFB_ACT_SWISH;   // networks.c:458 / :546-554
        // on the tensor path the last convolution writes the fp16 hi/lo planes the first input GEMM reads
        // (and the fp32 copy only when the caller wants to look at it)
        const bool planes = lastc && tc_gemm;
        if (lastc && c->use_tc_conv3) {
            // the convolution itself on the tensor cores (im2col view of the previous convolution's planes) ...
            LAUNCH(ffb_launch_conv_gemm_tc(c->d_c2hi.p, c->d_c2lo.p, (int64_t)m->conv_stride[i] * m->conv_nf[i], m->d_c3_hi, m->d_c3_lo,
                                           m->d_convb[i], keep ? out : nullptr, c->d_ahi.p, c->d_alo.p, Tt, m->conv_nfilter[i], m->c3_kp, st));
            // ... and the columns it cannot know about -- windows reaching into a neighbouring read, the reference's
            // right-edge plan -- once more on the CUDA cores from the fp32 input
            LAUNCH(ffb_launch_conv(cur, keep ? out : nullptr, c->d_ahi.p, c->d_alo.p, m->d_convWt[i], m->d_convb[i],
                                   c->d_geom[i].as<ffb::ReadGeom>(), c->d_tails[i].as<ffb::ConvTail>(), (int)N, c->col_off[i + 1][N],
                                   c->max_T[i + 1], m->conv_nf[i], m->conv_nfilter[i], m->conv_winlen[i], m->conv_stride[i], act,
                                   c->conv3_fix, st));
        } else {
            void *yhi = planes ? c->d_ahi.p : nullptr, *ylo = planes ? c->d_alo.p : nullptr;
            float *y = (planes && !keep) ? nullptr : out;
            if (i + 2 == m->nconv && c->use_tc_conv3) {
                // input of the tensor-core convolution: fp32 (for its fix-up pass) AND zero-initialised planes in the slot layout
                const size_t front = (size_t)((m->conv_winlen[i + 1] - 1) / 2) * m->conv_nf[i + 1];
                if (cudaMemsetAsync(c->d_c2hi.p, 0, c->d_c2hi.cap, st) != cudaSuccess || cudaMemsetAsync(c->d_c2lo.p, 0, c->d_c2lo.cap, st) != cudaSuccess)
                    return FFB_ERR_CUDA;
                yhi = (uint16_t *)c->d_c2hi.p + front; ylo = (uint16_t *)c->d_c2lo.p + front;
                y = out;
            }
            LAUNCH(ffb_launch_conv(cur, y, yhi, ylo, m->d_convWt[i], m->d_convb[i], c->d_geom[i].as<ffb::ReadGeom>(),
                                   c->d_tails[i].as<ffb::ConvTail>(), (int)N, c->col_off[i + 1][N], c->max_T[i + 1],
                                   m->conv_nf[i], m->conv_nfilter[i], m->conv_winlen[i], m->conv_stride[i], act, 0, st));
        }
        cur = out;
    }
    c->last_conv = c->d_act[0].as<float>();
    if (timed || kProfileBuild) cudaEventRecord(c->ev[1], st);
    // ---- five recurrent layers, directions B,F,B,F,B (networks.c:460-483 / :557-580) ----
    RnnBatch rb{c->d_order.as<int32_t>(), c->d_blkoff.as<int64_t>(), c->n_slots, (int)N};
    float gemm_ms = 0.f, rnn_ms = 0.f;
    c->ff_done = false;
    const bool tc_rnn = c->use_tc_rnn && tc_gemm;
    const bool tc_ff = tc_rnn && m->tc_ff && getenv("FFB_NO_TC_FF") == nullptr;
    const float *in = c->d_act[0].as<float>();   // fp32 input of the current layer (NULL when only planes exist)
    // streamed mode: GEMM l+1 is launched behind recurrence l and eats its output planes as they appear
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, m->device);
    const int free_sms = sm_count - c->tc_clusters * ffb_rnn_tc_cluster_size(m->kind, m->S);
    // GRU: layer l's recurrence writes the first S columns (z gate) of layer l+1's Xin itself; the GEMM does the rest
    const bool fuse_z = m->fuse_z && tc_rnn;
    const int n0_next = fuse_z ? S : 0;
    const bool streamed = c->stream_gemm && tc_rnn && !timed && free_sms >= (G * S - n0_next) / 128 && (G * S) / 128 <= 16;
    const size_t prog_stride = (size_t)c->n_groups + 1 + 16;   // per layer: group counters, finished-CTA counter, 16 ticket queues
    if (streamed) {
        if (cudaMemsetAsync(c->d_progress.p, 0, sizeof(int) * FFB_NLAYER * prog_stride, st) != cudaSuccess) return FFB_ERR_CUDA;
    }
    // ONE Xin buffer, also when streamed: the input GEMM of layer l+1 overwrites the projection of layer l IN PLACE.  A tile is
    // released to it only once layer l's recurrence has published the steps that produce the tile's rows -- by then it has
    // also read Xin for those rows (row t is read at step t and at no other time), and its own fused z columns of a row are
    // written by the very thread that read them.  Halves the largest workspace (3 KB per block): twice the reads per batch.
    float *xin_buf[2] = {c->d_xin.as<float>(), c->d_xin.as<float>()};
    for (int l = 0; l < FFB_NLAYER; l++) {
        const bool last = (l == FFB_NLAYER - 1);
        float *out = keep ? c->d_keep[l].as<float>() : c->d_act[1].as<float>();
        float *xin = xin_buf[l & 1];
        if (timed) cudaEventRecord(c->ev[5], st);
        if (tc_gemm) {
            if (!streamed || l == 0) {
                // layer 0 gets its fp16 hi/lo planes from the convolution, later layers from the tensor recurrent kernel
                if (l > 0 && !tc_rnn) LAUNCH(ffb_launch_split_f16(in, c->d_ahi.p, c->d_alo.p, Tt * m->layer_in[l], st));
                LAUNCH(ffb_launch_gemm_tc(c->d_ahi.p, c->d_alo.p, m->d_iW_hi[l], m->d_iW_lo[l], m->d_b[l], xin, Tt,
                                          G * S, m->layer_in[l], l > 0 ? n0_next : 0, st));
            }
        } else {
            LAUNCH(ffb_launch_sgemm_bias(in, m->d_iWt[l], m->d_b[l], xin, Tt, G * S, m->layer_in[l], st));
        }
        if (timed) cudaEventRecord(c->ev[6], st);
        // from here on the SMs beside the recurrence are free: let the next batch's head in.  Only a batch whose recurrence
        // fills most of the chip gates its successors -- small batches (the per-read drop-ins) gain from plain concurrency.
        if (last && !timed && tc_rnn && 2 * (sm_count - free_sms) >= sm_count) head_gate_publish(c);
        if (tc_rnn) {
            // the top layer feeds the output layer: fp16 planes for the tensor version, fp32 otherwise
            const bool ff_here = last && fuse_z && m->fuse_ff && tc_ff;     // the top layer writes trans itself
            const bool planes_out = !last || (tc_ff && !ff_here);
            float *out_f32 = (keep || (last && !tc_ff)) ? out : nullptr;
            int *prog = (streamed && !last) ? c->d_progress.as<int>() + (size_t)l * prog_stride : nullptr;
            RnnTcSched sched{c->d_slotoff.as<int32_t>(), c->d_slotlist.as<int32_t>(), c->tc_clusters, c->n_slots / 16};
            // (unstreamed, xin_buf[0] == xin_buf[1]: in place -- the thread that writes column j of a row has read it already)
            const bool fz = fuse_z && !last;
            const float ffs = m->head ? c->temperature : c->temperature / 5.0f;
            LAUNCH(ffb_launch_rnn_tc(m->kind, S, xin, m->d_sW_img[l], out_f32, planes_out ? c->d_ahi.p : nullptr,
                                     planes_out ? c->d_alo.p : nullptr, rb, sched, c->R_tc, (l % 2) == 0, c->d_ring.p, prog,
                                     fz ? m->d_b[l + 1] : (ff_here ? m->d_ffb_s : nullptr),
                                     fz ? xin_buf[(l + 1) & 1] : (ff_here ? c->d_trans.as<float>() : nullptr),
                                     ff_here ? nr : 0, ffs, st));
            c->ff_done = ff_here;
            if (streamed && !last) {
                const int dir = (l % 2) == 0 ? 1 : 0;   // layer l runs backward for even l (networks.c:460-483)
                LAUNCH(ffb_launch_gemm_tc_streamed(c->d_ahi.p, c->d_alo.p, m->d_iW_hi[l + 1], m->d_iW_lo[l + 1], m->d_b[l + 1],
                                                   xin_buf[(l + 1) & 1], Tt, G * S, S, c->d_work[dir].as<GemmWork>(), prog,
                                                   prog + c->n_groups + 1, n0_next, st));
            }
        } else {
            // the fp32 kernel ping-pongs between the two activation buffers
            if (!keep) out = c->d_act[(l & 1) ^ 1].as<float>();
            LAUNCH(ffb_launch_rnn(m->kind, S, xin, m->d_sWp[l], out, rb, (l % 2) == 0, st));
        }
        if (timed) {
            cudaEventRecord(c->ev[7], st);
            cudaEventSynchronize(c->ev[7]);
            float t1 = 0, t2 = 0;
            cudaEventElapsedTime(&t1, c->ev[5], c->ev[6]);
            cudaEventElapsedTime(&t2, c->ev[6], c->ev[7]);
            gemm_ms += t1; rnn_ms += t2;
        }
        in = out;
    }
    const float *top = in;
    if (timed || kProfileBuild) cudaEventRecord(c->ev[2], st);
    // ---- globalnorm_flipflop (layers.c:1082-1106) ----
    // scale: flip-flop divides tanh by temperature / 5 (layers.c:1087); the run-length head computes 5 tanhf / temperature
    const float ff_scale = m->head ? c->temperature : c->temperature / 5.0f;
    if (c->ff_done) {
        // trans was written by the top recurrent layer (its free MMA rows carry FF_W)
    } else if (tc_ff)
        LAUNCH(ffb_launch_ff_tanh_tc(c->d_ahi.p, c->d_alo.p, m->d_ff_hi, m->d_ff_lo, m->d_ffb_pad, c->d_trans.as<float>(), Tt, nr, S,
                                     ff_scale, m->head, st));
    else
        LAUNCH(ffb_launch_ff_tanh(top, m->d_ffWt, m->d_ffb, c->d_trans.as<float>(), Tt, nr, S, ff_scale, m->head, st));
    // -logZ/T (layers.c:1035-1096) shifts every transition score of a read by one constant.  The flip-flop posteriors
    // (running-shifted scans, per-block normalisation), their Viterbi path, qualities and trace are invariant under it,
    // so in forward-backward mode the fp64 partition scan only runs when the caller wants `trans`.  The run-length
    // posteriors are UNNORMALISED sums (decode.c:1096-1114): without the shift their fp32 forward values grow by
    // ~logZ/T per block and the decoder loses its low bits, so that head always normalises.
    const bool need_logz = m->head || (c->flags & FFB_FLAG_VITERBI_ONLY) || (c->flags & FFB_FLAG_WANT_TRANS) || getenv("FFB_ALWAYS_LOGZ");
    if (need_logz) {
        if (m->head) {
            LAUNCH(ffb_launch_rle_logz(c->d_trans.as<float>(), c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_logz.as<double>(), st));
        } else {
            LAUNCH(ffb_launch_logz(c->d_trans.as<float>(), c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_logz.as<double>(), st));
            LAUNCH(ffb_launch_sub_logz(c->d_trans.as<float>(), c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_logz.as<double>(), Tt, st));
        }
    }
    if (timed || kProfileBuild) cudaEventRecord(c->ev[3], st);
    // ---- decoding (flappie.c:277-300 / runnie.c:271-277) ----
    const float *post = c->d_trans.as<float>();
    if (m->head) {
        if (!(c->flags & FFB_FLAG_VITERBI_ONLY)) {
            LAUNCH(ffb_launch_rle_transpost(c->d_trans.as<float>(), c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_fwd.as<float>(),
                                            c->d_tpost.as<float>(), st));
            post = c->d_tpost.as<float>();
        }
        LAUNCH(ffb_launch_rle_viterbi(post, c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_tb.as<uint32_t>(), c->d_path.as<int32_t>(),
                                      c->d_qpath.as<float>(), c->d_score.as<float>(), st));
        LAUNCH(ffb_launch_rle_pack(c->d_trans.as<float>(), c->d_rle.as<float>(), Tt, st));
    } else {
        if (!(c->flags & FFB_FLAG_VITERBI_ONLY)) {
            LAUNCH(ffb_launch_transpost(c->d_trans.as<float>(), c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_fwd.as<float>(),
                                        c->d_tpost.as<float>(), Tt, st));   // includes the per-block log normalisation
            post = c->d_tpost.as<float>();
        }
        LAUNCH(ffb_launch_viterbi(post, c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_tb.as<uint64_t>(), c->d_path.as<int32_t>(),
                                  c->d_qpath.as<float>(), c->d_score.as<float>(), st));
        if (c->want_emit)    // change_positions + base / quality characters (flappie.c:284-297), --reverse included
            LAUNCH(ffb_launch_emit(c->d_path.as<int32_t>(), c->d_qpath.as<float>(), c->d_blkoff.as<int64_t>(), (int)N, m->nbase,
                                   (c->flags & FFB_FLAG_REVERSE) ? 1 : 0, m->d_phred_thr, m->n_phred_thr, c->d_bases.as<char>(),
                                   c->d_quals.as<char>(), c->d_nbases.as<int32_t>(), st));
        if (c->flags & FFB_FLAG_WANT_TRACE)
            LAUNCH(ffb_launch_trace(post, c->d_blkoff.as<int64_t>(), (int)N, nr, c->d_trace.p, 1, 0, st));
    }
    if (timed) {
        cudaEventRecord(c->ev[4], st);
        cudaEventSynchronize(c->ev[4]);
        c->t_gemm_ms = gemm_ms;
        c->t_rnn_ms = rnn_ms;
    } else if (kProfileBuild) {
        cudaEventRecord(c->ev[4], st);
    }
    return FFB_OK;
}

// profile build only: the group times of the LAST ordinary ffb_forward (streamed schedule, nothing synchronised in between):
// ms[0] convolutions, ms[1] input GEMM of layer 1 .. end of the recurrent layers, ms[2] output layer, ms[3] decoding
extern "C" int ffb_debug_group_times(ffb_ctx *c, float ms[4]) {
    if (!c || !ms || !kProfileBuild) return FFB_ERR_UNSUPPORTED;
    if (cudaStreamSynchronize(c->st) != cudaSuccess) return FFB_ERR_CUDA;
    for (int i = 0; i < 4; i++)
        if (cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); return FFB_ERR_CUDA; }
    return FFB_OK;
}

extern "C" int ffb_forward(ffb_ctx *c) {
    if (!c) return FFB_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->m->device), FFB_ERR_CUDA);
    return forward_impl(c, false);
}

extern "C" int ffb_forward_timed(ffb_ctx *c, float ms[8]) {
    if (!c || !ms) return FFB_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->m->device), FFB_ERR_CUDA);
    for (int i = 0; i < 8; i++) ms[i] = 0.f;
    const int r = forward_impl(c, true);
    if (r != FFB_OK) return r;
    if (c->n_reads == 0 || c->total_blocks == 0) return FFB_OK;
    float conv = 0, rec = 0, outl = 0, dec = 0, tot = 0;
    cudaEventElapsedTime(&conv, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&rec, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&outl, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&dec, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&tot, c->ev[0], c->ev[4]);
    ms[0] = conv; ms[1] = c->t_gemm_ms; ms[2] = c->t_rnn_ms; ms[3] = outl; ms[4] = dec; ms[5] = tot; ms[6] = rec;
    return FFB_OK;
}

// D2H of the requested outputs, enqueued on the context stream (no wait)
static int download_enqueue(ffb_ctx *c, const ffb_batch *b) {
    ffb_model *m = c->m;
    const int64_t N = c->n_reads, Tt = c->total_blocks;
    cudaStream_t st = c->st;
    if (N > 0 && Tt > 0) {
        if (b->path) CUDA_TRY(cudaMemcpyAsync(b->path, c->d_path.p, sizeof(int32_t) * (size_t)(Tt + N), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        if (b->qpath) CUDA_TRY(cudaMemcpyAsync(b->qpath, c->d_qpath.p, sizeof(float) * (size_t)(Tt + N), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        if (b->score) CUDA_TRY(cudaMemcpyAsync(b->score, c->d_score.p, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        if (c->want_emit && b->bases && b->quals && b->nbases) {
            CUDA_TRY(cudaMemcpyAsync(b->bases, c->d_bases.p, (size_t)(Tt + N), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
            CUDA_TRY(cudaMemcpyAsync(b->quals, c->d_quals.p, (size_t)(Tt + N), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
            CUDA_TRY(cudaMemcpyAsync(b->nbases, c->d_nbases.p, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        }
        if (b->trans && (c->flags & FFB_FLAG_WANT_TRANS))
            CUDA_TRY(cudaMemcpyAsync(b->trans, c->d_trans.p, sizeof(float) * (size_t)(Tt * m->nparam), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        if (b->tpost && (c->flags & FFB_FLAG_WANT_TRANS) && !(c->flags & FFB_FLAG_VITERBI_ONLY))
            CUDA_TRY(cudaMemcpyAsync(b->tpost, c->d_tpost.p, sizeof(float) * (size_t)(Tt * m->nparam), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        if (b->trace && (c->flags & FFB_FLAG_WANT_TRACE) && !m->head)
            CUDA_TRY(cudaMemcpyAsync(b->trace, c->d_trace.p, (size_t)((Tt + N) * m->nstate), cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
        if (b->rle_params && m->head)
            CUDA_TRY(cudaMemcpyAsync(b->rle_params, c->d_rle.p, sizeof(float) * 8 * (size_t)Tt, cudaMemcpyDeviceToHost, st), FFB_ERR_CUDA);
    }
    return FFB_OK;
}

// wait for everything enqueued on the context stream; rejected-only batches get their NAN scores here
extern "C" int ffb_collect(ffb_ctx *c, const ffb_batch *b) {
    if (!c || !b) return FFB_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->m->device), FFB_ERR_CUDA);
    CUDA_TRY(cudaStreamSynchronize(c->st), FFB_ERR_CUDA);
    if (b->score && c->total_blocks == 0)
        for (int64_t n = 0; n < c->n_reads; n++) b->score[n] = NAN;
    return FFB_OK;
}

extern "C" int ffb_download(ffb_ctx *c, const ffb_batch *b) {
    if (!c || !b) return FFB_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->m->device), FFB_ERR_CUDA);
    const int r = download_enqueue(c, b);
    if (r != FFB_OK) return r;
    return ffb_collect(c, b);
}

// submit = upload + all kernels + D2H enqueued, WITHOUT waiting for the results: with two contexts (two streams) the
// host prepares and uploads batch i+1 while the device still works on batch i; ffb_collect() waits for one of them
extern "C" int ffb_submit_batch(ffb_ctx *c, const ffb_batch *b) {
    int r = ffb_upload(c, b);
    if (r != FFB_OK) return r;
    r = ffb_forward(c);
    if (r != FFB_OK) return r;
    return download_enqueue(c, b);
}
extern "C" int ffb_submit_raw_batch(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b) {
    int r = ffb_upload_raw(c, rb, b);
    if (r != FFB_OK) return r;
    r = ffb_forward(c);
    if (r != FFB_OK) return r;
    return download_enqueue(c, b);
}

// the same in two halves (see upload_raw_begin): nothing is waited for in `begin`; `finish` blocks only until the trim bounds
// of THIS batch are back (a few hundred microseconds of device work, queued behind whatever else the device is doing)
extern "C" int ffb_submit_raw_begin(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b) { return upload_raw_begin(c, rb, b); }
extern "C" int ffb_submit_raw_finish(ffb_ctx *c) {
    int r = upload_raw_finish(c);
    if (r != FFB_OK) return r;
    r = ffb_forward(c);
    if (r != FFB_OK) return r;
    return download_enqueue(c, &c->raw_b);
}

extern "C" int ffb_basecall_batch(ffb_ctx *c, const ffb_batch *b) {
    int r = ffb_upload(c, b);
    if (r != FFB_OK) return r;
    r = ffb_forward(c);
    if (r != FFB_OK) return r;
    return ffb_download(c, b);
}

// pinned host memory for callers without the CUDA runtime (the C command line): async copies need it
extern "C" void *ffb_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    memset(p, 0, bytes);
    return p;
}
extern "C" void ffb_free_pinned(void *p) { if (p) cudaFreeHost(p); }

extern "C" int ffb_basecall_raw_batch(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b) {
    int r = ffb_upload_raw(c, rb, b);
    if (r != FFB_OK) return r;
    r = ffb_forward(c);
    if (r != FFB_OK) return r;
    return ffb_download(c, b);
}

extern "C" int64_t ffb_debug_fetch(ffb_ctx *c, int what, void *dst, int64_t bytes) {
    if (!c || !dst) return FFB_ERR_ARG;
    ffb_model *m = c->m;
    CUDA_TRY(cudaSetDevice(m->device), FFB_ERR_CUDA);
    const void *src = nullptr;
    int64_t have = 0;
    const int64_t Tt = c->total_blocks;
    if (what == 0) {
        src = c->last_conv; have = Tt * m->S * (int64_t)sizeof(float);
        if (!(c->flags & FFB_FLAG_KEEP_LAYERS)) { set_err("ffb_debug_fetch: conv output needs FFB_FLAG_KEEP_LAYERS"); return FFB_ERR_ARG; }
    } else if (what >= 1 && what <= FFB_NLAYER) {
        if (!(c->flags & FFB_FLAG_KEEP_LAYERS)) { set_err("ffb_debug_fetch: layers need FFB_FLAG_KEEP_LAYERS"); return FFB_ERR_ARG; }
        src = c->d_keep[what - 1].p; have = Tt * m->S * (int64_t)sizeof(float);
    } else if (what == 6) {
        src = c->d_trans.p; have = Tt * m->nparam * (int64_t)sizeof(float);
    } else if (what == 7) {
        src = c->d_logz.p; have = c->n_reads * (int64_t)sizeof(double);
    } else if (what == 8) {
        src = c->d_sig.p; have = c->total_samples * (int64_t)sizeof(float);
    } else {
        return FFB_ERR_ARG;
    }
    const int64_t n = std::min(have, bytes);
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)n, cudaMemcpyDeviceToHost, c->st), FFB_ERR_CUDA);
        CUDA_TRY(cudaStreamSynchronize(c->st), FFB_ERR_CUDA);
    }
    return n;
}

// ---- host-side base emission (flappie.c:284-297) -------------------------------------------
static inline char phredf_host(float p) {
    // qscoref / phredf, util.h:285-305
    const float p_clip = (p < 0.99999) ? p : 0.99999;
    const float q = -(10.0f * 0.43429448190325182765) * log1pf(-p_clip);
    char ph = roundf(33.0f + q);
    return (ph < 126) ? ph : 126;
}
static inline char phred_of_qpath(float x) { return phredf_host(expf(x)); }

// The quality character as a step function of qpath: thr[k] = the smallest float x with phred_of_qpath(x) >= 34 + k,
// found by bisection over the float ordering with THIS host's expf / log1pf.  The device emission kernel (emit.cu) only
// compares against the table, so its characters equal ffb_emit_bases' bit for bit.
static inline int32_t float_key(float f) { int32_t i; memcpy(&i, &f, 4); return i >= 0 ? i : (int32_t)(0x80000000u - (uint32_t)i); }
static inline float key_float(int32_t k) { int32_t i = k >= 0 ? k : (int32_t)(0x80000000u - (uint32_t)k); float f; memcpy(&f, &i, 4); return f; }
static const std::vector<float> &phred_thresholds() {
    static std::vector<float> thr;
    static std::once_flag once;
    std::call_once(once, [] {
        const int top = phred_of_qpath(88.0f);                 // expf saturates well before: the clipped maximum (83)
        for (int ch = 34; ch <= top; ch++) {
            int64_t lo = float_key(-1000.0f), hi = float_key(88.0f);   // f(lo) = 33 < ch <= f(hi)
            while (hi - lo > 1) {
                const int64_t mid = lo + (hi - lo) / 2;
                if (phred_of_qpath(key_float((int32_t)mid)) >= ch) hi = mid; else lo = mid;
            }
            thr.push_back(key_float((int32_t)hi));
        }
    });
    return thr;
}
extern "C" int ffb_phred_table(float *out, int cap) {
    const std::vector<float> &t = phred_thresholds();
    if (out) for (int k = 0; k < (int)t.size() && k < cap; k++) out[k] = t[(size_t)k];
    return (int)t.size();
}

extern "C" int ffb_emit_bases(const int32_t *path, const float *qpath, int64_t nblock, int nbase, bool reverse,
                              char *basecall, char *quality) {
    static const char lookup[5] = {'A', 'C', 'G', 'T', 'Z'};   // decode.h:16
    if (!path || !qpath || !basecall || !quality || nbase < 1 || nbase > 5) return -1;
    int n = 0;
    for (int64_t pos = 1; pos < nblock; pos++) {               // change_positions, decode.c:66-79 (npos = nblock)
        if (path[pos] == path[pos - 1]) continue;
        basecall[n] = lookup[path[pos] % nbase];
        quality[n] = phredf_host(expf(qpath[pos]));
        n++;
    }
    basecall[n] = 0; quality[n] = 0;
    if (reverse) {                                             // reverse_char_array, util.c:416
        std::reverse(basecall, basecall + n);
        std::reverse(quality, quality + n);
    }
    return n;
}

// runnie's run loop (runnie.c:279-310): one run per block whose state is a move state (< nbase); shape / scale are taken
// from the block where the run starts
extern "C" int64_t ffb_emit_runs(const int32_t *path, const float *rle_params, int64_t nblock, int nbase, char *bases,
                                 float *shape, float *scale, int32_t *dwell) {
    static const char lookup[5] = {'A', 'C', 'G', 'T', 'Z'};
    if (!path || !rle_params || !bases || !shape || !scale || !dwell || nbase != 4) return -1;
    int64_t n = 0, last_blk = -1;
    int32_t run = 1;
    auto emit = [&]() {
        const int base = path[last_blk];
        bases[n] = lookup[base];
        shape[n] = rle_params[last_blk * 2 * nbase + base];
        scale[n] = rle_params[last_blk * 2 * nbase + nbase + base];
        dwell[n] = run;
        n++;
    };
    for (int64_t blk = 0; blk < nblock; blk++) {
        if (path[blk] >= nbase) { run += 1; continue; }
        if (last_blk >= 0) emit();
        last_blk = blk;
        run = 1;
    }
    if (last_blk >= 0) emit();
    bases[n] = 0;
    return n;
}

// ======================================================================================
// (1) drop-ins with the reference's signatures
// ======================================================================================
extern "C" size_t nbase_from_flipflop_nparam(size_t nparam) {
    return (size_t)roundf((-1.0f + sqrtf(1 + 2 * nparam)) / 2.0f);   // layers.c:1029-1032
}

extern "C" size_t nbase_from_crf_runlength_nparam(size_t nparam) { return nbase_from_flipflop_nparam(nparam); }   // layers.c:1235-1239

// decode.c:66-79 -- the one other function of the replaced decode.c that the reference's callers use (flappie.c:284)
extern "C" size_t change_positions(int const *path, size_t npos, int *chpos) {
    if (!path || !chpos) return 0;
    size_t nch = 0;
    for (size_t pos = 1; pos < npos; pos++) {
        if (path[pos] == path[pos - 1]) continue;
        chpos[nch++] = (int)pos;
    }
    return nch;
}

extern "C" flappie_matrix make_flappie_matrix(size_t nr, size_t nc) {
    if (nr == 0 || nc == 0) return nullptr;
    const size_t nrq = (nr + 3) / 4;
    flappie_matrix mat = (flappie_matrix)malloc(sizeof(*mat));
    if (!mat) return nullptr;
    mat->nr = nr; mat->nrq = nrq; mat->nc = nc; mat->stride = nrq * 4;
    void *p = nullptr;
    if (posix_memalign(&p, 16, nrq * nc * 16) != 0) { free(mat); return nullptr; }
    memset(p, 0, nrq * nc * 16);
    mat->data.v = p;
    return mat;
}
extern "C" flappie_matrix free_flappie_matrix(flappie_matrix mat) {
    if (mat) { free(mat->data.v); free(mat); }
    return nullptr;
}
extern "C" flappie_imatrix make_flappie_imatrix(size_t nr, size_t nc) {
    return reinterpret_cast<flappie_imatrix>(make_flappie_matrix(nr, nc));
}
extern "C" flappie_imatrix free_flappie_imatrix(flappie_imatrix mat) {
    if (mat) { free(mat->data.v); free(mat); }
    return nullptr;
}

// flappie_matrix.c:342-358 (used by fast5_interface.c:write_trace): dense column-major copy, caller frees
extern "C" int32_t *array_from_flappie_imatrix(const_flappie_imatrix mat) {
    if (!mat) return nullptr;
    int32_t *res = (int32_t *)calloc(mat->nr * mat->nc, sizeof(int32_t));
    if (!res) return nullptr;
    for (size_t c = 0; c < mat->nc; c++)
        for (size_t r = 0; r < mat->nr; r++) res[c * mat->nr + r] = mat->data.f[c * mat->stride + r];
    return res;
}

extern "C" enum model_type get_flappie_model_type(const char *modelstr) {
    if (!modelstr) return FLAPPIE_MODEL_INVALID;
    if (0 == strcmp(modelstr, "r941_native")) return FLAPPIE_MODEL_R941_NATIVE;
    if (0 == strcmp(modelstr, "r941_rna002")) return FLAPPIE_MODEL_R941_RNA002;
    if (0 == strcmp(modelstr, "r941_5mC")) return FLAPPIE_MODEL_R941_5mC;
    if (0 == strcmp(modelstr, "r103_native")) return FLAPPIE_MODEL_R103_NATIVE;
    if (0 == strcmp(modelstr, "rle_r941_native")) return RUNNIE_MODEL_R941_NATIVE;
    // flappie <= 1.x name still used by the README and by BASELINE.json; the registry slot is
    // shared with r941_native (bind the GRU bundle to it with ffb_register_model)
    if (0 == strcmp(modelstr, "r10C_pcr")) return FLAPPIE_MODEL_R941_NATIVE;
    return FLAPPIE_MODEL_INVALID;
}
extern "C" const char *flappie_model_string(const enum model_type model) {
    switch (model) {
    case FLAPPIE_MODEL_R941_NATIVE: return "r941_native";
    case FLAPPIE_MODEL_R941_RNA002: return "r941_rna002";
    case FLAPPIE_MODEL_R941_5mC: return "r941_5mC";
    case FLAPPIE_MODEL_R103_NATIVE: return "r103_native";
    case RUNNIE_MODEL_R941_NATIVE: return "rle_r941_native";
    default: break;
    }
    fprintf(stderr, "Invalid model  %s:%d\n", __FILE__, __LINE__);   // reference: errx(EXIT_FAILURE, ...)
    exit(EXIT_FAILURE);
}
extern "C" const char *flappie_model_description(const enum model_type model) {
    switch (model) {
    case FLAPPIE_MODEL_R941_NATIVE: return "R9.4.1 model for MinION.  Trained from native DNA library";
    case FLAPPIE_MODEL_R941_RNA002: return "R9.4.1 dRNA model for MinION.  Trained from native and synthetic RNA library";
    case FLAPPIE_MODEL_R941_5mC: return "R9.4.1 model for PromethION; 5mC aware.  Trained from native NA12878 library";
    case FLAPPIE_MODEL_R103_NATIVE: return "R10.3 model for MinION.  Trained from native DNA library";
    case RUNNIE_MODEL_R941_NATIVE: return "R9.4.1 run-length encoded model for MinION.  Trained from native DNA library";
    default: break;
    }
    fprintf(stderr, "Invalid Flappie model  %s:%d\n", __FILE__, __LINE__);
    exit(EXIT_FAILURE);
}

// model registry for calculate_transitions().  The reference calls it from an OpenMP parallel-for (one read per call,
// flappie.c:364-385), so the lock covers only the registry: every caller takes a context of its own out of a pool (one
// stream + workspaces each), runs the read without the lock and puts the context back.
static std::mutex g_reg_mu;
static ffb_model *g_reg_model[RUNNIE_MODEL_INVALID + 1] = {nullptr};
static std::vector<ffb_ctx *> g_reg_pool[RUNNIE_MODEL_INVALID + 1];
static uint64_t g_reg_gen[RUNNIE_MODEL_INVALID + 1] = {0};

// Weight bundle file -> model (the format flappie_b200/host/ffb_host.h documents: "FFBW1", kind, nconv, strides, nmat,
// then nmat x {nr, nc, padded column-major floats} in guppy_model / guppy_stride5_model order).  The reference compiles
// its weights in (src/models/*.mdl, git-LFS); a process that only knows the reference API -- the reference's own main()
// linked against this library -- gets them from $FLAPPIE_B200_MODELS/<model name>.ffbw on first use.
extern "C" ffb_model *ffb_model_load(const char *path, int device) {
    if (!path) { set_err("ffb_model_load: NULL path"); return nullptr; }
    FILE *fp = fopen(path, "rb");
    if (!fp) { set_err("ffb_model_load: cannot open %s", path); return nullptr; }
    char magic[8];
    int32_t head[6];
    std::vector<_Mat> mats;
    std::vector<std::vector<float>> data;
    bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "FFBW1\0\0\0", 8) == 0 && fread(head, sizeof(int32_t), 6, fp) == 6 &&
              head[5] >= 1 && head[5] <= 64 && head[1] >= 1 && head[1] <= 3;
    for (int i = 0; ok && i < head[5]; i++) {
        uint64_t dim[2];
        ok = fread(dim, sizeof(uint64_t), 2, fp) == 2 && dim[0] > 0 && dim[1] > 0 && dim[0] <= (1u << 24) && dim[1] <= (1u << 24);
        if (!ok) break;
        _Mat m;
        m.nr = (size_t)dim[0]; m.nc = (size_t)dim[1]; m.nrq = (m.nr + 3) / 4; m.stride = 4 * m.nrq;
        data.emplace_back(m.stride * m.nc);
        ok = fread(data.back().data(), sizeof(float), data.back().size(), fp) == data.back().size();
        m.data.f = data.back().data();
        mats.push_back(m);
    }
    fclose(fp);
    if (!ok) { set_err("ffb_model_load: %s is not a weight bundle", path); return nullptr; }
    std::vector<const _Mat *> ptr;
    for (auto &m : mats) ptr.push_back(&m);
    const int stride[3] = {head[2], head[3], head[4]};
    return ffb_model_create(device, head[0], ptr.data(), (int)ptr.size(), stride, head[1]);
}

extern "C" int ffb_register_model(enum model_type which, ffb_model *m) {
    if ((int)which < 0 || (int)which >= RUNNIE_MODEL_INVALID || which == FLAPPIE_MODEL_INVALID) return FFB_ERR_ARG;
    std::vector<ffb_ctx *> old;
    {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        old.swap(g_reg_pool[which]);
        g_reg_model[which] = m;
        g_reg_gen[which]++;          // contexts still in use belong to the previous binding: destroyed when they come back
    }
    for (ffb_ctx *c : old) ffb_destroy(c);
    return FFB_OK;
}

extern "C" flappie_matrix calculate_transitions(const raw_table signal, float temperature, enum model_type model) {
    if (0 == signal.n || nullptr == signal.raw) return nullptr;          // networks.c:451-452
    if ((int)model < 0 || (int)model >= RUNNIE_MODEL_INVALID || model == FLAPPIE_MODEL_INVALID) {
        fprintf(stderr, "Invalid Flappie model  %s:%d\n", __FILE__, __LINE__);   // networks.c:98-104
        exit(EXIT_FAILURE);
    }
    ffb_model *m = nullptr;
    ffb_ctx *c = nullptr;
    uint64_t gen = 0;
    {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        m = g_reg_model[model];
        gen = g_reg_gen[model];
        if (m && !g_reg_pool[model].empty()) { c = g_reg_pool[model].back(); g_reg_pool[model].pop_back(); }
    }
    if (!m) {
        // nothing bound to this enum value yet: a caller that only knows the reference API
        const char *dir = getenv("FLAPPIE_B200_MODELS");
        if (dir) {
            static std::mutex load_mu;
            std::lock_guard<std::mutex> lk(load_mu);
            { std::lock_guard<std::mutex> lk2(g_reg_mu); m = g_reg_model[model]; gen = g_reg_gen[model]; }
            if (!m) {
                const std::string path = std::string(dir) + "/" + flappie_model_string(model) + ".ffbw";
                int dev = 0;
                if (getenv("FLAPPIE_B200_DEVICE")) dev = atoi(getenv("FLAPPIE_B200_DEVICE"));
                m = ffb_model_load(path.c_str(), dev);
                if (m) { ffb_register_model(model, m); std::lock_guard<std::mutex> lk2(g_reg_mu); gen = g_reg_gen[model]; }
            }
        }
    }
    if (!m) { set_err("calculate_transitions: no weights registered for model %d (ffb_register_model, or $FLAPPIE_B200_MODELS/<name>.ffbw)", (int)model); return nullptr; }
    if (!c) c = ffb_create(m, nullptr);
    if (!c) return nullptr;
    // the context goes back to the pool on every path out of this function
    struct Return {
        ffb_ctx *c; int which; uint64_t gen;
        ~Return() {
            bool keep = false;
            {
                std::lock_guard<std::mutex> lk(g_reg_mu);
                if (g_reg_gen[which] == gen) { g_reg_pool[which].push_back(c); keep = true; }
            }
            if (!keep) ffb_destroy(c);
        }
    } give_back{c, (int)model, gen};
    if (signal.end <= signal.start) return nullptr;
    const int64_t n = (int64_t)(signal.end - signal.start);
    const long T = ffb_model_nblock(m, n);
    if (T <= 0) return nullptr;
    int64_t off[2] = {0, n};
    std::vector<float> dense((size_t)T * m->nparam);
    ffb_batch b;
    memset(&b, 0, sizeof b);
    b.signal = signal.raw + signal.start; b.sig_off = off; b.n_reads = 1; b.temperature = temperature;
    b.flags = FFB_FLAG_VITERBI_ONLY | FFB_FLAG_WANT_TRANS;
    b.trans = dense.data();
    if (ffb_basecall_batch(c, &b) != FFB_OK) return nullptr;
    flappie_matrix out = make_flappie_matrix(m->nparam, (size_t)T);
    if (!out) return nullptr;
    for (long t = 0; t < T; t++) memcpy(out->data.f + (size_t)t * out->stride, dense.data() + (size_t)t * m->nparam, sizeof(float) * m->nparam);
    return out;
}

// standalone decode contexts (no model needed): one per device, made on the device that is current at the first call
// from it; each drop-in makes that device current again and holds the context's own lock while it uses the buffers
struct DecodeCtx {
    int device = 0;
    cudaStream_t st = nullptr;
    std::mutex mu;
    DevBuf trans, tpost, fwd, tb, path, qpath, score, blkoff, trace;
};
static DecodeCtx *decode_ctx() {
    static std::map<int, DecodeCtx *> ctxs;
    static std::mutex mu;
    if (ffb_device_count() < 1) { set_err("no CUDA device: the flappie_b200 decode entry points have no CPU fallback"); return nullptr; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_err("cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lk(mu);
    auto it = ctxs.find(dev);
    if (it != ctxs.end()) return it->second;
    DecodeCtx *d = new DecodeCtx();
    d->device = dev;
    if (cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete d; return nullptr; }
    ctxs[dev] = d;
    return d;
}
#define DECODE_LOCK(d) std::lock_guard<std::mutex> lk((d)->mu); cudaSetDevice((d)->device)

// _Mat [nr x nc] (padded columns) -> device dense [nc][nr]
static int mat_to_device(const _Mat *mat, DevBuf &buf, cudaStream_t st) {
    if (buf.reserve(sizeof(float) * mat->nr * mat->nc) != 0) return -1;
    if (cudaMemcpy2DAsync(buf.p, mat->nr * sizeof(float), mat->data.f, mat->stride * sizeof(float), mat->nr * sizeof(float),
                          mat->nc, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
    return 0;
}

extern "C" float decode_crf_flipflop(const_flappie_matrix trans, bool combine_stays, int *path, float *qpath) {
    if (!trans || !path || !qpath) return NAN;                           // decode.c:120-122
    const int nr = (int)trans->nr;
    const int64_t T = (int64_t)trans->nc;
    if (nr != 40 && nr != 60) { set_err("decode_crf_flipflop: unsupported nr=%d", nr); return NAN; }
    DecodeCtx *d = decode_ctx();
    if (!d) return NAN;
    DECODE_LOCK(d);
    int64_t off[2] = {0, T};
    bool ok = mat_to_device(trans, d->trans, d->st) == 0;
    ok = ok && d->tb.reserve(sizeof(uint64_t) * (size_t)T) == 0 && d->path.reserve(sizeof(int32_t) * (size_t)(T + 1)) == 0 &&
         d->qpath.reserve(sizeof(float) * (size_t)(T + 1)) == 0 && d->score.reserve(sizeof(float)) == 0 &&
         d->blkoff.reserve(sizeof(off)) == 0;
    if (!ok) { set_err("decode_crf_flipflop: device allocation / copy failed"); return NAN; }
    cudaMemcpyAsync(d->blkoff.p, off, sizeof off, cudaMemcpyHostToDevice, d->st);
    if (ffb_launch_viterbi(d->trans.as<float>(), d->blkoff.as<int64_t>(), 1, nr, d->tb.as<uint64_t>(), d->path.as<int32_t>(),
                           d->qpath.as<float>(), d->score.as<float>(), d->st) < 0) return NAN;
    float score = NAN;
    static_assert(sizeof(int) == sizeof(int32_t), "int is 32-bit");
    cudaMemcpyAsync(path, d->path.p, sizeof(int32_t) * (size_t)(T + 1), cudaMemcpyDeviceToHost, d->st);
    cudaMemcpyAsync(qpath, d->qpath.p, sizeof(float) * (size_t)(T + 1), cudaMemcpyDeviceToHost, d->st);
    cudaMemcpyAsync(&score, d->score.p, sizeof(float), cudaMemcpyDeviceToHost, d->st);
    if (cudaStreamSynchronize(d->st) != cudaSuccess) { set_err("decode_crf_flipflop: %s", cudaGetErrorString(cudaGetLastError())); return NAN; }
    if (combine_stays) {                                                 // decode.c:194-198
        const int nbase = (int)nbase_from_flipflop_nparam(nr);
        for (int64_t i = 0; i <= T; i++) path[i] = (path[i] < nbase) ? path[i] : -1;
    }
    return score;
}

extern "C" flappie_matrix transpost_crf_flipflop(const_flappie_matrix trans, bool return_log) {
    if (!trans) return nullptr;
    const int nr = (int)trans->nr;
    const int64_t T = (int64_t)trans->nc;
    if (nr != 40 && nr != 60) { set_err("transpost_crf_flipflop: unsupported nr=%d", nr); return nullptr; }
    DecodeCtx *d = decode_ctx();
    if (!d) return nullptr;
    DECODE_LOCK(d);
    const int nstate = 2 * (int)nbase_from_flipflop_nparam(nr);
    int64_t off[2] = {0, T};
    bool ok = mat_to_device(trans, d->trans, d->st) == 0;
    ok = ok && d->tpost.reserve(sizeof(float) * (size_t)(T * nr)) == 0 && d->fwd.reserve(2 * sizeof(float) * (size_t)((T + 1) * nstate)) == 0 &&
         d->blkoff.reserve(sizeof(off)) == 0;
    if (!ok) { set_err("transpost_crf_flipflop: device allocation / copy failed"); return nullptr; }
    cudaMemcpyAsync(d->blkoff.p, off, sizeof off, cudaMemcpyHostToDevice, d->st);
    if (ffb_launch_transpost(d->trans.as<float>(), d->blkoff.as<int64_t>(), 1, nr, d->fwd.as<float>(), d->tpost.as<float>(), T, d->st) < 0) return nullptr;
    if (!return_log && ffb_launch_exp_inplace(d->tpost.as<float>(), T * nr, d->st) < 0) return nullptr;
    flappie_matrix out = make_flappie_matrix(nr, (size_t)T);
    if (!out) return nullptr;
    cudaMemcpy2DAsync(out->data.f, out->stride * sizeof(float), d->tpost.p, nr * sizeof(float), nr * sizeof(float), (size_t)T,
                      cudaMemcpyDeviceToHost, d->st);
    if (cudaStreamSynchronize(d->st) != cudaSuccess) { set_err("transpost_crf_flipflop: %s", cudaGetErrorString(cudaGetLastError())); return free_flappie_matrix(out); }
    return out;
}

// ---- run-length drop-ins (decode.h:48-49) ----
extern "C" float decode_crf_runlength(const_flappie_matrix param, int *path) {
    if (!param || !path) return NAN;                                     // decode.c:902-903
    const int nr = (int)param->nr;
    const int64_t T = (int64_t)param->nc;
    if (nr != 40) { set_err("decode_crf_runlength: unsupported nr=%d", nr); return NAN; }
    DecodeCtx *d = decode_ctx();
    if (!d) return NAN;
    DECODE_LOCK(d);
    int64_t off[2] = {0, T};
    bool ok = mat_to_device(param, d->trans, d->st) == 0;
    ok = ok && d->tb.reserve(sizeof(uint64_t) * (size_t)std::max<int64_t>(T, 1)) == 0 && d->path.reserve(sizeof(int32_t) * (size_t)(T + 1)) == 0 &&
         d->qpath.reserve(sizeof(float) * (size_t)(T + 1)) == 0 && d->score.reserve(sizeof(float)) == 0 && d->blkoff.reserve(sizeof(off)) == 0;
    if (!ok) { set_err("decode_crf_runlength: device allocation / copy failed"); return NAN; }
    cudaMemcpyAsync(d->blkoff.p, off, sizeof off, cudaMemcpyHostToDevice, d->st);
    if (ffb_launch_rle_viterbi(d->trans.as<float>(), d->blkoff.as<int64_t>(), 1, nr, d->tb.as<uint32_t>(), d->path.as<int32_t>(),
                               d->qpath.as<float>(), d->score.as<float>(), d->st) < 0) return NAN;
    float score = NAN;
    cudaMemcpyAsync(path, d->path.p, sizeof(int32_t) * (size_t)T, cudaMemcpyDeviceToHost, d->st);
    cudaMemcpyAsync(&score, d->score.p, sizeof(float), cudaMemcpyDeviceToHost, d->st);
    if (cudaStreamSynchronize(d->st) != cudaSuccess) { set_err("decode_crf_runlength: %s", cudaGetErrorString(cudaGetLastError())); return NAN; }
    return score;
}

extern "C" flappie_matrix transpost_crf_runlength(const_flappie_matrix param) {
    if (!param) return nullptr;
    const int nr = (int)param->nr;
    const int64_t T = (int64_t)param->nc;
    if (nr != 40) { set_err("transpost_crf_runlength: unsupported nr=%d", nr); return nullptr; }
    DecodeCtx *d = decode_ctx();
    if (!d) return nullptr;
    DECODE_LOCK(d);
    int64_t off[2] = {0, T};
    bool ok = mat_to_device(param, d->trans, d->st) == 0;
    ok = ok && d->tpost.reserve(sizeof(float) * (size_t)(T * nr)) == 0 && d->fwd.reserve(2 * sizeof(float) * (size_t)((T + 1) * 8)) == 0 &&
         d->blkoff.reserve(sizeof(off)) == 0;
    if (!ok) { set_err("transpost_crf_runlength: device allocation / copy failed"); return nullptr; }
    cudaMemcpyAsync(d->blkoff.p, off, sizeof off, cudaMemcpyHostToDevice, d->st);
    if (ffb_launch_rle_transpost(d->trans.as<float>(), d->blkoff.as<int64_t>(), 1, nr, d->fwd.as<float>(), d->tpost.as<float>(), d->st) < 0) return nullptr;
    flappie_matrix out = make_flappie_matrix(nr, (size_t)T);
    if (!out) return nullptr;
    cudaMemcpy2DAsync(out->data.f, out->stride * sizeof(float), d->tpost.p, nr * sizeof(float), nr * sizeof(float), (size_t)T,
                      cudaMemcpyDeviceToHost, d->st);
    if (cudaStreamSynchronize(d->st) != cudaSuccess) { set_err("transpost_crf_runlength: %s", cudaGetErrorString(cudaGetLastError())); return free_flappie_matrix(out); }
    return out;
}

extern "C" void exp_activation_inplace(flappie_matrix C) {
    if (!C) return;
    DecodeCtx *d = decode_ctx();
    if (!d) return;
    DECODE_LOCK(d);
    const size_t n = C->stride * C->nc;   // the reference exponentiates the padding too (layers.c:58-63)
    if (d->trans.reserve(sizeof(float) * n) != 0) { set_err("exp_activation_inplace: device allocation failed"); return; }
    cudaMemcpyAsync(d->trans.p, C->data.f, sizeof(float) * n, cudaMemcpyHostToDevice, d->st);
    ffb_launch_exp_inplace(d->trans.as<float>(), (int64_t)n, d->st);
    cudaMemcpyAsync(C->data.f, d->trans.p, sizeof(float) * n, cudaMemcpyDeviceToHost, d->st);
    if (cudaStreamSynchronize(d->st) != cudaSuccess) set_err("exp_activation_inplace: %s", cudaGetErrorString(cudaGetLastError()));
}

extern "C" flappie_imatrix trace_from_posterior(flappie_matrix tpost) {
    if (!tpost) return nullptr;
    const int nr = (int)tpost->nr;
    const int64_t T = (int64_t)tpost->nc;
    if (nr != 40 && nr != 60) { set_err("trace_from_posterior: unsupported nr=%d", nr); return nullptr; }
    DecodeCtx *d = decode_ctx();
    if (!d) return nullptr;
    DECODE_LOCK(d);
    const int nstate = 2 * (int)nbase_from_flipflop_nparam(nr);
    int64_t off[2] = {0, T};
    bool ok = mat_to_device(tpost, d->trans, d->st) == 0 && d->trace.reserve(sizeof(int32_t) * (size_t)((T + 1) * nstate)) == 0 &&
              d->blkoff.reserve(sizeof(off)) == 0;
    if (!ok) { set_err("trace_from_posterior: device allocation / copy failed"); return nullptr; }
    cudaMemcpyAsync(d->blkoff.p, off, sizeof off, cudaMemcpyHostToDevice, d->st);
    // int32 entries: the reference keeps round(255 * sum) as an int (decode.c:520-540), 256 included
    if (ffb_launch_trace(d->trans.as<float>(), d->blkoff.as<int64_t>(), 1, nr, d->trace.p, 0, 1, d->st) < 0) return nullptr;
    std::vector<int32_t> h((size_t)((T + 1) * nstate));
    cudaMemcpyAsync(h.data(), d->trace.p, sizeof(int32_t) * h.size(), cudaMemcpyDeviceToHost, d->st);
    if (cudaStreamSynchronize(d->st) != cudaSuccess) { set_err("trace_from_posterior: %s", cudaGetErrorString(cudaGetLastError())); return nullptr; }
    flappie_imatrix out = make_flappie_imatrix(nstate, (size_t)(T + 1));
    if (!out) return nullptr;
    for (int64_t t = 0; t <= T; t++)
        for (int s = 0; s < nstate; s++) out->data.f[(size_t)t * out->stride + s] = h[(size_t)t * nstate + s];
    return out;
}
