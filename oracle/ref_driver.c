/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver around the UNMODIFIED reference sources compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/libflappie_ref.so.
 * The reference's networks.c cannot be compiled (its model headers are git-LFS
 * pointers), so this file re-states ONLY the layer call
 * order of
 *     flipflop_guppy_transitions   (reference src/networks.c:450-489)  GRU topology
 *     flipflop5_guppy_transitions  (reference src/networks.c:539-586)  LSTM topology
 * and the post-network half of calculate_post (reference src/flappie.c:262-300)
 * around a weight bundle handed in at run time.  Every arithmetic routine that
 * runs below is the reference's own object code.
 */
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#include "decode.h"
#include "flappie_output.h"
#include "flappie_matrix.h"
#include "flappie_structures.h"
#include "layers.h"
#include "nnfeatures.h"
#include "util.h"

#define FFREF_MAX_CONV 3
#define FFREF_NLAYER 5

/* kind: 0 = GRU topology (guppy_model, networks.c:150-177),
 *       1 = LSTM topology (guppy_stride5_model, networks.c:180-215). */
typedef struct {
    int kind;
    int nconv;
    flappie_matrix conv_W[FFREF_MAX_CONV];
    flappie_matrix conv_b[FFREF_MAX_CONV];
    int conv_stride[FFREF_MAX_CONV];
    flappie_matrix iW[FFREF_NLAYER];
    flappie_matrix sW[FFREF_NLAYER];
    flappie_matrix b[FFREF_NLAYER];
    flappie_matrix FF_W;
    flappie_matrix FF_b;
} ffref_model;

/* Build a reference _Mat from dense column-major data (nr x nc, ld = nr). */
flappie_matrix ffref_mat_from_dense(const float *x, size_t nr, size_t nc) {
    return mat_from_array(x, nr, nc);
}

/* Copy a reference _Mat into dense column-major data. */
void ffref_mat_to_dense(const_flappie_matrix m, float *out) {
    for (size_t c = 0; c < m->nc; c++) {
        memcpy(out + c * m->nr, m->data.f + c * m->stride, m->nr * sizeof(float));
    }
}

void ffref_imat_to_dense(const_flappie_imatrix m, int32_t *out) {
    for (size_t c = 0; c < m->nc; c++) {
        memcpy(out + c * m->nr, m->data.f + c * m->stride, m->nr * sizeof(int32_t));
    }
}

size_t ffref_mat_nr(const_flappie_matrix m) { return m->nr; }
size_t ffref_mat_nc(const_flappie_matrix m) { return m->nc; }

/* Network forward.  If `dump` is non-NULL it must have room for 8 matrices and
 * receives copies of: [0..nconv-1] conv outputs (post activation),
 * [3..7] the five recurrent layer outputs.  Caller frees them. */
flappie_matrix ffref_transitions(const ffref_model *net, const float *signal, size_t n,
                                 float temperature, flappie_matrix *dump) {
    raw_table rt = {.uuid = NULL, .n = n, .start = 0, .end = n, .raw = (float *)signal};
    flappie_matrix cur = features_from_raw(rt);
    for (int i = 0; i < net->nconv; i++) {
        flappie_matrix nxt = convolution(cur, net->conv_W[i], net->conv_b[i], net->conv_stride[i], NULL);
        if (net->kind == 0) {
            tanh_activation_inplace(nxt);   /* networks.c:458 */
        } else {
            swish_activation_inplace(nxt);  /* networks.c:546,550,554 */
        }
        cur = free_flappie_matrix(cur);
        cur = nxt;
        if (dump) dump[i] = copy_flappie_matrix(cur);
    }
    for (int l = 0; l < FFREF_NLAYER; l++) {
        flappie_matrix xin = feedforward_linear(cur, net->iW[l], net->b[l], NULL);
        cur = free_flappie_matrix(cur);
        const bool backward = (l % 2) == 0;   /* B,F,B,F,B  networks.c:460-483 / 557-580 */
        if (net->kind == 0) {
            cur = backward ? grumod_backward(xin, net->sW[l], NULL) : grumod_forward(xin, net->sW[l], NULL);
        } else {
            cur = backward ? lstm_backward(xin, net->sW[l], NULL) : lstm_forward(xin, net->sW[l], NULL);
        }
        xin = free_flappie_matrix(xin);
        if (dump) dump[3 + l] = copy_flappie_matrix(cur);
    }
    flappie_matrix trans = globalnorm_flipflop(cur, net->FF_W, net->FF_b, temperature, NULL);
    cur = free_flappie_matrix(cur);
    return trans;
}

/* Post-network half of calculate_post (reference src/flappie.c:262-300).
 * path/qpath need nblock+2 entries, basecall/quality nblock+1 chars,
 * trace (if non-NULL) nstate*(nblock+1) int32.  Returns number of bases or -1. */
long ffref_decode(flappie_matrix trans_weights, bool viterbi_only, int *path, float *qpath,
                  char *basecall, char *quality, float *score_out, int32_t *trace_out,
                  float *post_out) {
    if (NULL == trans_weights) return -1;
    const size_t nbase = nbase_from_flipflop_nparam(trans_weights->nr);
    const size_t nblock = trans_weights->nc;
    int *path_idx = calloc(nblock + 2, sizeof(int));
    flappie_matrix posterior = trans_weights;
    if (!viterbi_only) {
        posterior = transpost_crf_flipflop(trans_weights, true);
    }
    if (post_out) ffref_mat_to_dense(posterior, post_out);
    const float score = decode_crf_flipflop(posterior, false, path, qpath);
    const size_t nidx = change_positions(path, nblock, path_idx);
    for (size_t i = 0; i < nidx; i++) {
        const size_t idx = path_idx[i];
        basecall[i] = base_lookup[path[idx] % nbase];
        quality[i] = phredf(expf(qpath[idx]));
    }
    basecall[nidx] = 0;
    quality[nidx] = 0;
    if (trace_out) {
        flappie_matrix pcopy = copy_flappie_matrix(posterior);
        exp_activation_inplace(pcopy);
        flappie_imatrix trace = trace_from_posterior(pcopy);
        ffref_imat_to_dense(trace, trace_out);
        trace = free_flappie_imatrix(trace);
        pcopy = free_flappie_matrix(pcopy);
    }
    if (posterior != trans_weights) posterior = free_flappie_matrix(posterior);
    free(path_idx);
    *score_out = score;
    return (long)nidx;
}

/* Whole hot path for one already-normalised read; used for CPU-baseline timing.
 * Returns number of bases called (or -1). */
long ffref_basecall(const ffref_model *net, const float *signal, size_t n, float temperature,
                    bool viterbi_only, char *basecall, char *quality, float *score_out) {
    flappie_matrix trans = ffref_transitions(net, signal, n, temperature, NULL);
    if (NULL == trans) return -1;
    const size_t nblock = trans->nc;
    int *path = calloc(nblock + 2, sizeof(int));
    float *qpath = calloc(nblock + 2, sizeof(float));
    int32_t *trace = malloc(sizeof(int32_t) * 2 * nbase_from_flipflop_nparam(trans->nr) * (nblock + 1));
    /* reference always computes the trace (flappie.c:299-300) */
    long nb = ffref_decode(trans, viterbi_only, path, qpath, basecall, quality, score_out, trace, NULL);
    free(trace);
    free(qpath);
    free(path);
    trans = free_flappie_matrix(trans);
    return nb;
}


/* One record through the reference's own writers (src/flappie_output.c:92-133) into the file `path` (appended):
 * fmt 0 = fasta, 1 = fastq, 2 = sam -- by NAME, so the enum order of flappie_output.h does not matter here. */
int ffref_format(const char *fmt_name, const char *path, const char *uuid, const char *readname, bool uuid_primary,
                 const char *prefix, float score, size_t nblock, const char *basecall, const char *quality,
                 size_t n, size_t start, size_t end) {
    const enum flappie_outformat_type fmt = get_outformat(fmt_name);
    if (FLAPPIE_OUTFORMAT_INVALID == fmt) return -1;
    FILE *fp = fopen(path, "a");
    if (NULL == fp) return -1;
    struct _raw_basecall_info res = {0};
    res.score = score;
    res.rt.n = n; res.rt.start = start; res.rt.end = end;
    res.basecall = (char *)basecall;
    res.quality = (char *)quality;
    res.basecall_length = strlen(basecall);
    res.nblock = nblock;
    fprintf_format(fmt, fp, uuid, readname, uuid_primary, prefix, res);
    fclose(fp);
    return 0;
}
