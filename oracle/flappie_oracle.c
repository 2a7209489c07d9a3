/* oracle/flappie_oracle.c -- TEST INFRASTRUCTURE ONLY (see flappie_oracle.h).
 *
 * Plain-C restatement of the reference algorithm for the hot path.  Written from the
 * reference's behaviour, function by function, each citing the file:line it follows.
 * No BLAS, no SSE: scalar fp32 in the reference's visit order wherever the order is
 * defined by the reference itself (Viterbi, forward/backward, partition function);
 * the SGEMM/SGEMV sums (whose order OpenBLAS does not define) are accumulated in
 * increasing-k order.
 */
#include "flappie_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- scalar math: reference src/util.h:276-305,331-339 ------------------------ */

/* logisticfv = 1 / (1 + exp(-x))  (util.h:331-334) */
static inline float ffo_logistic(float x) { return 1.0f / (1.0f + expf(-x)); }
/* tanhfv = 2 * logistic(2x) - 1   (util.h:336-339) */
static inline float ffo_tanh(float x) {
    const float y = ffo_logistic(x + x);
    return (y + y) - 1.0f;
}
/* util.h:276-278 */
static inline float ffo_logsumexpf(float x, float y) {
    return fmaxf(x, y) + log1pf(expf(-fabsf(x - y)));
}
/* util.h:280-282 */
static inline double ffo_logsumexp(double x, double y) {
    return fmax(x, y) + log1p(exp(-fabs(x - y)));
}

static inline int iceil_(int x, int y) { return (x + y - 1) / y; }

static float ffo_activate(float x, int act) {
    switch (act) {
    case FFO_ACT_TANH: return ffo_tanh(x);          /* layers.c:40-48 */
    case FFO_ACT_SWISH: return x * ffo_logistic(x); /* layers.c:24-32 */
    default: return x;
    }
}

/* ---- convolution: reference src/layers.c:189-276 ------------------------------ */

long ffo_conv_plan(int T, int winlen, int stride, ffo_conv_term *terms, long cap, int *ncol_out) {
    if (T < winlen || winlen < 1 || stride < 1) return -1;
    const long padL = (winlen - 1) / 2;           /* layers.c:202 */
    const long padR = winlen / 2;                 /* layers.c:203 */
    const long ncolC = iceil_(T, stride);         /* layers.c:204 */
    if (ncol_out) *ncol_out = (int)ncolC;
    long n = 0;
#define EMIT(c, xs, tl, nt)                                                         \
    do {                                                                            \
        const long c_ = (c), xs_ = (xs), tl_ = (tl), nt_ = (nt);                    \
        if (nt_ > 0) {                                                              \
            if (c_ < 0 || c_ >= ncolC || xs_ < 0 || xs_ + nt_ > T || tl_ < 0 ||     \
                tl_ + nt_ > winlen)                                                 \
                return -1;                                                          \
            if (n < cap) terms[n] = (ffo_conv_term){(int32_t)c_, (int32_t)xs_, (int32_t)tl_, (int32_t)nt_}; \
            n++;                                                                    \
        }                                                                           \
    } while (0)

    /* Left edge: only part of the filter covers the input (layers.c:219-226) */
    for (long w = 0; w < padL; w += stride) {
        EMIT(w / stride, 0, padL - w, winlen - (padL - w));
    }
    const long ncolsL_complete = iceil_((int)padL, stride);      /* layers.c:229 */
    const long shiftX_L = ncolsL_complete * stride - padL;       /* layers.c:233 */
    const long nstepC = iceil_(winlen, stride);                  /* layers.c:236 */
    const long nstepX = stride * nstepC;                         /* layers.c:237 */
    /* Interior: one strided SGEMM per phase (layers.c:239-254) */
    for (long w = 0; w < winlen; w += stride) {
        const long ncol_processed = (T - shiftX_L - w) / nstepX; /* ifloor, layers.c:248 */
        const long initial_col = w / stride;
        for (long i = 0; i < ncol_processed; i++) {
            EMIT(ncolsL_complete + initial_col + i * nstepC, shiftX_L + w + i * nstepX, 0, winlen);
        }
    }
    /* Right edge (layers.c:257-271) */
    const long maxCol_reshape = (T - shiftX_L) / nstepX;
    const long remainder_reshape = (T - shiftX_L) % nstepX;
    const long colR = ncolsL_complete + nstepC * (maxCol_reshape - 1) + remainder_reshape / stride + 1;
    const long xR = T - winlen + 1;
    const long startR = stride - (padL + T - winlen) % stride - 1;
    for (long w = startR; w < padR; w += stride) {
        EMIT(colR + w / stride, xR + w, 0, winlen - (w + 1));
    }
#undef EMIT
    return n;
}

int ffo_convolution(const float *x, int T, int nf, const float *W, const float *b, int nfilter,
                    int winlen, int stride, int act, float *out) {
    int ncol = 0;
    long nterm = ffo_conv_plan(T, winlen, stride, NULL, 0, &ncol);
    if (nterm < 0) return -1;
    ffo_conv_term *terms = malloc(sizeof(ffo_conv_term) * (size_t)(nterm > 0 ? nterm : 1));
    if (!terms) return -1;
    ffo_conv_plan(T, winlen, stride, terms, nterm, &ncol);
    /* bias (layers.c:215-217) */
    for (int c = 0; c < ncol; c++) memcpy(out + (size_t)c * nfilter, b, sizeof(float) * nfilter);
    for (long t = 0; t < nterm; t++) {
        const ffo_conv_term q = terms[t];
        float *o = out + (size_t)q.col * nfilter;
        for (int f = 0; f < nfilter; f++) {
            const float *w = W + ((size_t)f * winlen + q.tap_lo) * nf;
            const float *xi = x + (size_t)q.x_start * nf;
            float acc = 0.0f;
            for (int j = 0; j < q.ntap * nf; j++) acc += w[j] * xi[j];
            o[f] += acc;
        }
    }
    free(terms);
    const size_t tot = (size_t)ncol * nfilter;
    for (size_t i = 0; i < tot; i++) out[i] = ffo_activate(out[i], act);
    return ncol;
}

/* ---- affine map: reference src/flappie_matrix.c:361-389 ------------------------ */

void ffo_affine(const float *X, int T, int K, const float *W, const float *b, int N, float *out) {
    /* transpose W to [K][N] so the inner loop runs over independent outputs */
    float *Wt = malloc(sizeof(float) * (size_t)K * N);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) Wt[(size_t)k * N + n] = W[(size_t)n * K + k];
    float *acc = malloc(sizeof(float) * (size_t)N);
    for (int t = 0; t < T; t++) {
        const float *x = X + (size_t)t * K;
        float *o = out + (size_t)t * N;
        for (int n = 0; n < N; n++) acc[n] = 0.0f;
        for (int k = 0; k < K; k++) {
            const float xv = x[k];
            const float *w = Wt + (size_t)k * N;
            for (int n = 0; n < N; n++) acc[n] += xv * w[n];
        }
        for (int n = 0; n < N; n++) o[n] = b[n] + acc[n];
    }
    free(acc);
    free(Wt);
}

/* ---- grumod: reference src/layers.c:571-715 ------------------------------------ */

void ffo_grumod(const float *Xin, int T, int S, const float *sW, int backward, float *h) {
    float *Wt = malloc(sizeof(float) * (size_t)S * 3 * S);   /* [k][3S] */
    for (int n = 0; n < 3 * S; n++)
        for (int k = 0; k < S; k++) Wt[(size_t)k * 3 * S + n] = sW[(size_t)n * S + k];
    float *a = malloc(sizeof(float) * 3 * (size_t)S);
    float *prev = calloc((size_t)S, sizeof(float));           /* zero initial state :592,:638 */
    for (int step = 0; step < T; step++) {
        const int t = backward ? (T - 1 - step) : step;       /* :640-652 / :597-607 */
        const float *x = Xin + (size_t)t * 3 * S;
        float *o = h + (size_t)t * S;
        /* grumod_step (layers.c:664-715) */
        for (int n = 0; n < 3 * S; n++) a[n] = 0.0f;
        for (int k = 0; k < S; k++) {
            const float hv = prev[k];
            const float *w = Wt + (size_t)k * 3 * S;
            for (int n = 0; n < 3 * S; n++) a[n] += hv * w[n];
        }
        for (int j = 0; j < S; j++) {
            const float z = ffo_logistic(x[j] + a[j]);             /* :690-699 */
            const float r = ffo_logistic(x[S + j] + a[S + j]);
            const float hbar = ffo_tanh(r * a[2 * S + j] + x[2 * S + j]);   /* :704-709 */
            o[j] = z * prev[j] + (1.0f - z) * hbar;                /* :712-714 */
        }
        memcpy(prev, o, sizeof(float) * (size_t)S);
    }
    free(prev);
    free(a);
    free(Wt);
}

/* ---- lstm: reference src/layers.c:877-1026 ------------------------------------- */

void ffo_lstm(const float *Xin, int T, int S, const float *sW, int backward, float *out) {
    float *Wt = malloc(sizeof(float) * (size_t)S * 4 * S);   /* [k][4S] */
    for (int n = 0; n < 4 * S; n++)
        for (int k = 0; k < S; k++) Wt[(size_t)k * 4 * S + n] = sW[(size_t)n * S + k];
    float *a = malloc(sizeof(float) * 4 * (size_t)S);
    float *prev = calloc((size_t)S, sizeof(float));           /* zero output :902,:953 */
    float *state = calloc((size_t)S, sizeof(float));          /* zero cell state :892 */
    for (int step = 0; step < T; step++) {
        const int t = backward ? (T - 1 - step) : step;
        const float *x = Xin + (size_t)t * 4 * S;
        float *o = out + (size_t)t * S;
        for (int n = 0; n < 4 * S; n++) a[n] = 0.0f;
        for (int k = 0; k < S; k++) {
            const float hv = prev[k];
            const float *w = Wt + (size_t)k * 4 * S;
            for (int n = 0; n < 4 * S; n++) a[n] += hv * w[n];
        }
        for (int j = 0; j < S; j++) {
            /* lstm_step (layers.c:1013-1025): chunks (i, f, g, o) */
            const float forget = ffo_logistic(x[S + j] + a[S + j]) * state[j];
            const float update = ffo_logistic(x[j] + a[j]) * ffo_tanh(x[2 * S + j] + a[2 * S + j]);
            state[j] = forget + update;
            o[j] = ffo_logistic(x[3 * S + j] + a[3 * S + j]) * ffo_tanh(state[j]);
        }
        memcpy(prev, o, sizeof(float) * (size_t)S);
    }
    free(state);
    free(prev);
    free(a);
    free(Wt);
}

/* ---- global normalisation: reference src/layers.c:1029-1106 -------------------- */

int ffo_nbase_from_nparam(int nparam) {
    /* layers.c:1029-1032 */
    return (int)roundf((-1.0f + sqrtf(1 + 2 * nparam)) / 2.0f);
}

static double ffo_partition(const float *C, int T, int nr) {
    /* crf_manystay_partition_function (layers.c:1035-1079) */
    const int nbase = ffo_nbase_from_nparam(nr);
    const int nstate = nbase + nbase;
    double mem[2 * 32] = {0};
    double *curr = mem, *prev = mem + nstate;
    for (int c = 0; c < T; c++) {
        const float *col = C + (size_t)c * nr;
        const float *stay = col + nstate * nbase;
        double *tmp = curr; curr = prev; prev = tmp;
        for (int s = nbase; s < nstate; s++) {
            const int from = s - nbase;
            curr[s] = ffo_logsumexp(prev[s] + stay[s], prev[from] + stay[from]);
        }
        for (int to = 0; to < nbase; to++) {
            const float *row = col + to * nstate;
            curr[to] = row[0] + prev[0];
            for (int from = 1; from < nstate; from++)
                curr[to] = ffo_logsumexp(curr[to], row[from] + prev[from]);
        }
    }
    double logZ = curr[0];
    for (int s = 1; s < nstate; s++) logZ = ffo_logsumexp(logZ, curr[s]);
    return logZ;
}

void ffo_globalnorm_flipflop(const float *h, int T, int S, const float *W, const float *b, int nr,
                             float temperature, float *trans, double *logZ_out) {
    /* globalnorm_manystay (layers.c:1082-1100) */
    ffo_affine(h, T, S, W, b, nr, trans);
    const size_t tot = (size_t)T * nr;
    const float scale = temperature / 5.0f;                 /* :1087 */
    for (size_t i = 0; i < tot; i++) trans[i] = (ffo_tanh(trans[i]) - 0.0f) / scale;   /* flappie_matrix.c:625-633 */
    const double logZd = ffo_partition(trans, T, nr);
    if (logZ_out) *logZ_out = logZd;
    const float logZ = (float)(logZd / (double)T);          /* :1089 */
    for (size_t i = 0; i < tot; i++) trans[i] -= logZ;
}

/* ---- Viterbi: reference src/decode.c:104-204 ----------------------------------- */

static inline int ffo_trans_lookup(int from, int to, int nbase) {
    /* decode.c:104-114 */
    const int nstate = nbase + nbase;
    return (to < nbase) ? (to * nstate + from) : (nbase * nstate + from);
}

float ffo_decode_crf_flipflop(const float *trans, int T, int nr, int *path, float *qpath) {
    const int nbase = ffo_nbase_from_nparam(nr);            /* decode.c:125 */
    const int nstate = nbase + nbase;
    float mem[2 * 32] = {0};                                /* calloc :130 */
    int *tb = malloc(sizeof(int) * (size_t)nstate * (size_t)(T > 0 ? T : 1));
    float *curr = mem, *prev = mem + nstate;
    for (int blk = 0; blk < T; blk++) {
        const float *col = trans + (size_t)blk * nr;
        const float *flop = col + nstate * nbase;
        int *tbc = tb + (size_t)blk * nstate;
        float *tmp = curr; curr = prev; prev = tmp;
        for (int b2 = nbase; b2 < nstate; b2++) {           /* :153-164 */
            curr[b2] = prev[b2] + flop[b2];
            tbc[b2] = b2;
            const int from = b2 - nbase;
            const float score = prev[from] + flop[from];
            if (score > curr[b2]) { curr[b2] = score; tbc[b2] = from; }
        }
        for (int b1 = 0; b1 < nbase; b1++) {                /* :167-180 */
            const float *row = col + b1 * nstate;
            curr[b1] = row[0] + prev[0];
            tbc[b1] = 0;
            for (int from = 1; from < nstate; from++) {
                const float score = row[from] + prev[from];
                if (score > curr[b1]) { curr[b1] = score; tbc[b1] = from; }
            }
        }
    }
    /* valmaxf / argmaxf: first max wins (util.c:17-31,47-61) */
    int imax = 0;
    float vmax = curr[0];
    for (int s = 1; s < nstate; s++) if (curr[s] > vmax) { vmax = curr[s]; imax = s; }
    path[T] = imax;
    for (int blk = T; blk > 0; blk--) {                     /* :186-191 */
        path[blk - 1] = tb[(size_t)(blk - 1) * nstate + path[blk]];
        qpath[blk] = trans[(size_t)(blk - 1) * nr + ffo_trans_lookup(path[blk - 1], path[blk], nbase)];
    }
    qpath[0] = NAN;                                         /* :192 */
    free(tb);
    return vmax;
}

/* ---- transition posteriors: reference src/decode.c:377-497 --------------------- */

int ffo_transpost_crf_flipflop(const float *trans, int T, int nr, int return_log, float *tpost) {
    const int nbase = ffo_nbase_from_nparam(nr);
    const int nstate = nbase + nbase;
    float *fwd = calloc((size_t)nstate * (size_t)(T + 1), sizeof(float));
    if (!fwd) return -1;
    for (int blk = 0; blk < T; blk++) {                     /* forwards :396-423 */
        const float *col = trans + (size_t)blk * nr;
        const float *flop = col + nstate * nbase;
        const float *prev = fwd + (size_t)blk * nstate;
        float *curr = fwd + (size_t)(blk + 1) * nstate;
        for (int b2 = nbase; b2 < nstate; b2++) {
            const int from = b2 - nbase;
            curr[b2] = ffo_logsumexpf(prev[b2] + flop[b2], prev[from] + flop[from]);
        }
        for (int b1 = 0; b1 < nbase; b1++) {
            const float *row = col + b1 * nstate;
            curr[b1] = row[0] + prev[0];
            for (int from = 1; from < nstate; from++)
                curr[b1] = ffo_logsumexpf(curr[b1], row[from] + prev[from]);
        }
    }
    float mem[2 * 32] = {0};
    float *prev = mem, *curr = mem + nstate;
    for (int blk = T; blk > 0; blk--) {                     /* backwards :434-484 */
        const float *f = fwd + (size_t)(blk - 1) * nstate;
        const float *col = trans + (size_t)(blk - 1) * nr;
        const float *flop = col + nstate * nbase;
        float *pcol = tpost + (size_t)(blk - 1) * nr;
        float *pflop = pcol + nstate * nbase;
        float *tmp = prev; prev = curr; curr = tmp;
        for (int b1 = 0; b1 < nbase; b1++)                  /* :446-453 */
            for (int st = 0; st < nstate; st++)
                pcol[b1 * nstate + st] = f[st] + prev[b1] + col[b1 * nstate + st];
        for (int b = nbase; b < nstate; b++) {              /* :454-461 */
            const int fb = b - nbase;
            pflop[b] = f[b] + prev[b] + flop[b];
            pflop[fb] = f[fb] + prev[b] + flop[fb];
        }
        for (int b2 = nbase; b2 < nstate; b2++) {           /* :465-471 */
            const int from = b2 - nbase;
            curr[b2] = prev[b2] + flop[b2];
            curr[from] = prev[b2] + flop[from];
        }
        for (int b1 = 0; b1 < nbase; b1++) {                /* :474-482 */
            const float *row = col + b1 * nstate;
            for (int from = 0; from < nstate; from++)
                curr[from] = ffo_logsumexpf(curr[from], row[from] + prev[b1]);
        }
    }
    free(fwd);
    /* log_row_normalise_inplace (flappie_matrix.c:450-467) */
    for (int blk = 0; blk < T; blk++) {
        float *pcol = tpost + (size_t)blk * nr;
        float lse = pcol[0];
        for (int r = 1; r < nr; r++) lse = ffo_logsumexpf(lse, pcol[r]);
        for (int r = 0; r < nr; r++) pcol[r] -= lse;
    }
    if (!return_log) {
        const size_t tot = (size_t)T * nr;
        for (size_t i = 0; i < tot; i++) tpost[i] = expf(tpost[i]);
    }
    return 0;
}

/* ---- trace: reference src/decode.c:499-543 ------------------------------------- */

void ffo_trace_from_posterior(const float *tpost, int T, int nr, int32_t *trace) {
    const int nbase = ffo_nbase_from_nparam(nr);
    const int nstate = nbase + nbase;
    for (int from = 0; from < nstate; from++) {             /* :511-518 */
        float sum = 0.0f;
        for (int to = 0; to < nbase; to++) sum += tpost[to * nstate + from];
        sum += tpost[nbase * nstate + from];
        trace[from] = (int32_t)roundf(255.0f * sum);
    }
    for (int blk = 0; blk < T; blk++) {                     /* :521-540 */
        int32_t *tr = trace + (size_t)(blk + 1) * nstate;
        const float *pcol = tpost + (size_t)blk * nr;
        for (int to = 0; to < nbase; to++) {
            const float *row = pcol + to * nstate;
            float sum = row[0];
            for (int from = 1; from < nstate; from++) sum += row[from];
            tr[to] = (int32_t)roundf(255.0f * sum);
        }
        const float *pflop = pcol + nbase * nstate;
        for (int to = nbase; to < nstate; to++) {
            const float sum = pflop[to - nbase] + pflop[to];
            tr[to] = (int32_t)roundf(255.0f * sum);
        }
    }
}

/* ---- base emission: reference src/decode.c:66-79, src/flappie.c:284-292 -------- */

int ffo_change_positions(const int *path, int npos, int *chpos) {
    int nch = 0;
    for (int pos = 1; pos < npos; pos++) {
        if (path[pos] == path[pos - 1]) continue;
        chpos[nch++] = pos;
    }
    return nch;
}

char ffo_phredf(float p) {
    /* qscoref / phredf (util.h:285-305) */
    const float p_clip = (p < 0.99999) ? p : 0.99999;
    const float q = -(10.0f * 0.43429448190325182765) * log1pf(-p_clip);
    char ph = roundf(33.0f + q);
    return (ph < 126) ? ph : 126;
}

int ffo_emit_bases(const int *path, const float *qpath, int nblock, int nbase, char *basecall,
                   char *quality) {
    static const char lookup[5] = {'A', 'C', 'G', 'T', 'Z'};   /* decode.h:16 */
    int *idx = malloc(sizeof(int) * (size_t)(nblock + 2));
    const int n = ffo_change_positions(path, nblock, idx);     /* flappie.c:284 */
    for (int i = 0; i < n; i++) {
        basecall[i] = lookup[path[idx[i]] % nbase];            /* :289 */
        quality[i] = ffo_phredf(expf(qpath[idx[i]]));          /* :290 */
    }
    basecall[n] = 0;
    quality[n] = 0;
    free(idx);
    return n;
}

/* ---- whole network: reference src/networks.c:450-489, 539-586 ------------------ */

int ffo_nblock(const ffo_model *m, int nsample) {
    int T = nsample;
    for (int i = 0; i < m->nconv; i++) {
        if (T < m->conv_winlen[i]) return -1;
        T = iceil_(T, m->conv_stride[i]);
    }
    return T;
}

int ffo_transitions(const ffo_model *m, const float *signal, int n, float temperature, float *trans,
                    float *conv_out, float **layers_out) {
    const int G = (m->kind == FFO_GRU) ? 3 : 4;
    const int S = m->size;
    int T = n;
    float *cur = malloc(sizeof(float) * (size_t)n);
    memcpy(cur, signal, sizeof(float) * (size_t)n);            /* features_from_raw nnfeatures.c:15 */
    int width = 1;
    for (int i = 0; i < m->nconv; i++) {
        const int To = iceil_(T, m->conv_stride[i]);
        float *nxt = malloc(sizeof(float) * (size_t)To * m->conv_nfilter[i]);
        const int act = (m->kind == FFO_GRU) ? FFO_ACT_TANH : FFO_ACT_SWISH;
        if (ffo_convolution(cur, T, m->conv_nf[i], m->conv_W[i], m->conv_b[i], m->conv_nfilter[i],
                            m->conv_winlen[i], m->conv_stride[i], act, nxt) < 0) {
            free(nxt);
            free(cur);
            return -1;
        }
        free(cur);
        cur = nxt;
        T = To;
        width = m->conv_nfilter[i];
    }
    if (conv_out) memcpy(conv_out, cur, sizeof(float) * (size_t)T * width);
    float *xin = malloc(sizeof(float) * (size_t)T * G * S);
    for (int l = 0; l < 5; l++) {
        ffo_affine(cur, T, width, m->iW[l], m->b[l], G * S, xin);
        float *nxt = malloc(sizeof(float) * (size_t)T * S);
        const int backward = (l % 2) == 0;                     /* B,F,B,F,B */
        if (m->kind == FFO_GRU) ffo_grumod(xin, T, S, m->sW[l], backward, nxt);
        else ffo_lstm(xin, T, S, m->sW[l], backward, nxt);
        free(cur);
        cur = nxt;
        width = S;
        if (layers_out && layers_out[l]) memcpy(layers_out[l], cur, sizeof(float) * (size_t)T * S);
    }
    free(xin);
    ffo_globalnorm_flipflop(cur, T, S, m->FF_W, m->FF_b, m->nparam, temperature, trans, NULL);
    free(cur);
    return T;
}

int ffo_basecall(const ffo_model *m, const float *signal, int n, float temperature, int viterbi_only,
                 char *basecall, char *quality, float *score, int *path, float *qpath,
                 int32_t *trace) {
    const int T = ffo_nblock(m, n);
    if (T < 0) return -1;
    const int nr = m->nparam;
    const int nbase = ffo_nbase_from_nparam(nr);
    float *trans = malloc(sizeof(float) * (size_t)T * nr);
    if (ffo_transitions(m, signal, n, temperature, trans, NULL, NULL) < 0) { free(trans); return -1; }
    float *post = trans;
    if (!viterbi_only) {                                       /* flappie.c:278-282 */
        post = malloc(sizeof(float) * (size_t)T * nr);
        ffo_transpost_crf_flipflop(trans, T, nr, 1, post);
    }
    int *p = path ? path : malloc(sizeof(int) * (size_t)(T + 2));
    float *q = qpath ? qpath : malloc(sizeof(float) * (size_t)(T + 2));
    *score = ffo_decode_crf_flipflop(post, T, nr, p, q);       /* :283 */
    const int nb = ffo_emit_bases(p, q, T, nbase, basecall, quality);
    if (trace) {                                               /* :299-300 */
        const size_t tot = (size_t)T * nr;
        float *e = malloc(sizeof(float) * tot);
        for (size_t i = 0; i < tot; i++) e[i] = expf(post[i]);
        ffo_trace_from_posterior(e, T, nr, trace);
        free(e);
    }
    if (!path) free(p);
    if (!qpath) free(q);
    if (post != trans) free(post);
    free(trans);
    return nb;
}


/* ==================================================================================
 * Run-length ("runnie") head: reference src/layers.c:1235-1358, src/decode.c:893-1159,
 * src/runnie.c:241-316.  Parameters per block (nbase = 4: 40 rows):
 *   [0, nbase)        shape of the run-length distribution, 1 + softplus(x)
 *   [nbase, 2 nbase)  scale, 1e-8 + softplus(x)
 *   [2 nbase, nr)     transition scores 5 tanh(x) / temperature - logZ / T, indexed by
 *                     rle_index(from, stay_from, to, stay_to).
 * States: b < nbase = "move into base b" (emits a base), b + nbase = "stay in base b".
 * ================================================================================== */

static inline int ffo_rle_index(int base_from, int stay_from, int base_to, int nbase) {
    /* decode.c:893-898 / layers.c:1241-1246: the destination's stay flag is implied by base_from == base_to */
    return base_to * 2 * nbase + base_from + (stay_from ? nbase : 0);
}
static inline float ffo_softplusf(float x) { return log1pf(expf(-fabsf(x))) + ((x >= 0.0f) ? x : 0.f); }   /* util.h:83-85 */
static inline double ffo_logsumexp_d(double x, double y) { return fmax(x, y) + log1p(exp(-fabs(x - y))); } /* util.h:272-274 */

double ffo_runlength_partition(const float *C, int T, int nr) {
    /* runlengthV2_partition_function, layers.c:1255-1302: double state, the stay states through the FLOAT
     * logsumexpf (its arguments and its result are rounded to float) */
    const int nbase = ffo_nbase_from_nparam(nr), nstate = 2 * nbase;
    double mem[2 * 16] = {0};
    double *curr = mem, *prev = mem + nstate;
    for (int c = 0; c < T; c++) {
        const float *p = C + (size_t)c * nr + nstate;
        double *tmp = curr; curr = prev; prev = tmp;
        for (int b1 = 0; b1 < nbase; b1++) {
            curr[b1] = -HUGE_VAL;
            for (int b2 = 0; b2 < nbase; b2++) {
                if (b1 == b2) continue;
                curr[b1] = ffo_logsumexp_d(curr[b1], prev[b2] + p[ffo_rle_index(b2, 0, b1, nbase)]);
                curr[b1] = ffo_logsumexp_d(curr[b1], prev[b2 + nbase] + p[ffo_rle_index(b2, 1, b1, nbase)]);
            }
        }
        for (int b = 0; b < nbase; b++) {
            const float x = (float)(prev[b] + p[ffo_rle_index(b, 0, b, nbase)]);
            const float y = (float)(prev[b + nbase] + p[ffo_rle_index(b, 1, b, nbase)]);
            curr[b + nbase] = ffo_logsumexpf(x, y);
        }
    }
    double logZ = curr[0];
    for (int st = 1; st < nstate; st++) logZ = ffo_logsumexp_d(logZ, curr[st]);
    return logZ;
}

void ffo_globalnorm_runlength(const float *h, int T, int S, const float *W, const float *b, int nr,
                              float temperature, float *C, double *logZ_out) {
    /* globalnorm_runlengthV2, layers.c:1326-1358 */
    ffo_affine(h, T, S, W, b, nr, C);
    const int nbase = ffo_nbase_from_nparam(nr), nrun = 2 * nbase;
    for (int c = 0; c < T; c++) {
        float *col = C + (size_t)c * nr;
        for (int k = 0; k < nbase; k++) {
            col[k] = 1.0f + ffo_softplusf(col[k]);
            col[nbase + k] = 1e-8f + ffo_softplusf(col[nbase + k]);
        }
        for (int r = nrun; r < nr; r++) col[r] = 5.0f * tanhf(col[r]) / temperature;
    }
    const double Z = ffo_runlength_partition(C, T, nr);
    if (logZ_out) *logZ_out = Z;
    const float logZ = (float)(Z / (float)T);               /* :1349: double / float, then rounded to float */
    for (int c = 0; c < T; c++)
        for (int r = nrun; r < nr; r++) C[(size_t)c * nr + r] -= logZ;
}

float ffo_decode_crf_runlength(const float *param, int T, int nr, int *path) {
    /* decode_crf_runlength, decode.c:901-984; path has T entries */
    const int nbase = ffo_nbase_from_nparam(nr), nstate = 2 * nbase;
    float mem[2 * 16] = {0};
    char *tb = calloc((size_t)nstate * (size_t)(T > 0 ? T : 1), 1);
    float *prev = mem, *curr = mem + nstate;
    for (int blk = 0; blk < T; blk++) {
        const float *p = param + (size_t)blk * nr + nstate;
        char *tbc = tb + (size_t)blk * nstate;
        float *tmp = prev; prev = curr; curr = tmp;
        for (int st = 0; st < nstate; st++) curr[st] = -HUGE_VALF;
        for (int b1 = 0; b1 < nbase; b1++) {
            for (int b2 = 0; b2 < nbase; b2++) {
                if (b1 == b2) continue;
                const float mv = prev[b2] + p[ffo_rle_index(b2, 0, b1, nbase)];
                if (mv > curr[b1]) { curr[b1] = mv; tbc[b1] = (char)b2; }
                const float sv = prev[b2 + nbase] + p[ffo_rle_index(b2, 1, b1, nbase)];
                if (sv > curr[b1]) { curr[b1] = sv; tbc[b1] = (char)(b2 + nbase); }
            }
        }
        for (int b = 0; b < nbase; b++) {
            const float sv = prev[b + nbase] + p[ffo_rle_index(b, 1, b, nbase)];
            const float mv = prev[b] + p[ffo_rle_index(b, 0, b, nbase)];
            if (sv > mv) { curr[b + nbase] = sv; tbc[b + nbase] = (char)(b + nbase); }
            else { curr[b + nbase] = mv; tbc[b + nbase] = (char)b; }
        }
    }
    int last = 0;                                           /* argmaxf: first max wins (util.c:17-31) */
    for (int st = 1; st < nstate; st++) if (curr[st] > curr[last]) last = st;
    const float score = curr[last];
    for (int blk = T; blk > 0; blk--) {
        const int from = tb[(size_t)(blk - 1) * nstate + last];
        path[blk - 1] = last;
        last = from;
    }
    free(tb);
    return score;
}

int ffo_transpost_crf_runlength(const float *param, int T, int nr, float *post) {
    /* transpost_crf_runlength, decode.c:1013-1159: UNNORMALISED log posteriors fwd + bwd + score; the shape and
     * scale rows are copied through */
    const int nbase = ffo_nbase_from_nparam(nr), nstate = 2 * nbase;
    float *fwd = calloc((size_t)nstate * (size_t)(T + 1), sizeof(float));
    if (!fwd) return -1;
    for (int blk = 0; blk < T; blk++) {
        const float *p = param + (size_t)blk * nr + nstate;
        const float *prev = fwd + (size_t)blk * nstate;
        float *curr = fwd + (size_t)(blk + 1) * nstate;
        for (int b1 = 0; b1 < nbase; b1++) {
            curr[b1] = -HUGE_VALF;
            for (int b2 = 0; b2 < nbase; b2++) {
                if (b1 == b2) continue;
                const float sv = prev[b2 + nbase] + p[ffo_rle_index(b2, 1, b1, nbase)];
                const float mv = prev[b2] + p[ffo_rle_index(b2, 0, b1, nbase)];
                curr[b1] = ffo_logsumexpf(curr[b1], ffo_logsumexpf(sv, mv));
            }
        }
        for (int b = 0; b < nbase; b++) {
            const float sv = prev[b + nbase] + p[ffo_rle_index(b, 1, b, nbase)];
            const float mv = prev[b] + p[ffo_rle_index(b, 0, b, nbase)];
            curr[b + nbase] = ffo_logsumexpf(sv, mv);
        }
    }
    float mem[2 * 16] = {0};
    float *prev = mem, *curr = mem + nstate;
    for (int blk = T; blk > 0; blk--) {
        const float *f = fwd + (size_t)(blk - 1) * nstate;
        const float *p = param + (size_t)(blk - 1) * nr + nstate;
        float *q = post + (size_t)(blk - 1) * nr + nstate;
        float *tmp = curr; curr = prev; prev = tmp;
        for (int b1 = 0; b1 < nbase; b1++) {
            curr[b1] = -HUGE_VALF;
            curr[b1 + nbase] = -HUGE_VALF;
            for (int b2 = 0; b2 < nbase; b2++) {
                if (b1 == b2) continue;
                const int mi = ffo_rle_index(b1, 0, b2, nbase);
                curr[b1] = ffo_logsumexpf(curr[b1], prev[b2] + p[mi]);
                q[mi] = f[b1] + prev[b2] + p[mi];
                const int si = ffo_rle_index(b1, 1, b2, nbase);
                curr[b1 + nbase] = ffo_logsumexpf(curr[b1 + nbase], prev[b2] + p[si]);
                q[si] = f[b1 + nbase] + prev[b2] + p[si];
            }
        }
        for (int b = 0; b < nbase; b++) {
            const int i = ffo_rle_index(b, 0, b, nbase);
            curr[b] = ffo_logsumexpf(curr[b], prev[b + nbase] + p[i]);
            q[i] = f[b] + p[i] + prev[b + nbase];
        }
        for (int b = 0; b < nbase; b++) {
            const int i = ffo_rle_index(b, 1, b, nbase);
            curr[b + nbase] = ffo_logsumexpf(curr[b + nbase], prev[b + nbase] + p[i]);
            q[i] = f[b + nbase] + p[i] + prev[b + nbase];
        }
        for (int k = 0; k < nstate; k++) post[(size_t)(blk - 1) * nr + k] = param[(size_t)(blk - 1) * nr + k];
    }
    free(fwd);
    return 0;
}

int ffo_emit_runs(const int *path, const float *post, int T, int nr, char *bases, float *shape, float *scale, int *dwell) {
    /* the run loop of runnie's calculate_post, runnie.c:279-310: one run per block whose state is a move state */
    static const char lookup[5] = {'A', 'C', 'G', 'T', 'Z'};
    const int nbase = ffo_nbase_from_nparam(nr);
    int n = 0, run = 1, last_blk = -1;
    for (int blk = 0; blk < T; blk++) {
        if (path[blk] >= nbase) { run += 1; continue; }
        if (last_blk >= 0) {
            const int base = path[last_blk];
            bases[n] = lookup[base]; shape[n] = post[(size_t)last_blk * nr + base];
            scale[n] = post[(size_t)last_blk * nr + nbase + base]; dwell[n] = run; n++;
        }
        last_blk = blk;
        run = 1;
    }
    if (last_blk >= 0) {
        const int base = path[last_blk];
        bases[n] = lookup[base]; shape[n] = post[(size_t)last_blk * nr + base];
        scale[n] = post[(size_t)last_blk * nr + nbase + base]; dwell[n] = run; n++;
    }
    return n;
}
