/* oracle/flappie_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the flappie hot path (network forward + flip-flop
 * decoding).  It is the CHECKER for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (flappie_b200/csrc) never links or calls anything declared here.
 *
 * Parity status: PINNED -- every function below is compared in tests/test_oracle.py
 * against the reference's own object code (oracle/_ref, built by oracle/Makefile from
 * /root/reference/src) and against golden vectors under tests/golden/ generated from
 * that object code by tests/golden/make_golden.py.
 *
 * Layouts are dense (no _Mat row padding): activations are [block][feature]
 * row-major; weight matrices are [out][in] row-major, which is byte-identical to the
 * reference's column-major W[nr=in, nc=out] with the padding stripped
 * (reference src/flappie_matrix.c:361-389: C = W^T X + b).
 */
#ifndef FLAPPIE_ORACLE_H
#define FLAPPIE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FFO_ACT_TANH = 0, FFO_ACT_SWISH = 1, FFO_ACT_NONE = 2 };
enum { FFO_GRU = 0, FFO_LSTM = 1 };

/* One contribution of the reference's convolution(): out[col] += sum_{j<ntap}
 * W[tap_lo+j] . x[x_start+j]  (reference src/layers.c:219-271). */
typedef struct {
    int32_t col, x_start, tap_lo, ntap;
} ffo_conv_term;

/* Integer plan of reference convolution() for T input columns.  Writes at most
 * `cap` terms; returns the number of terms the reference issues (may exceed cap),
 * or -1 if the reference's own index arithmetic would leave the matrices (T too
 * small).  *ncol_out = iceil(T, stride). */
long ffo_conv_plan(int T, int winlen, int stride, ffo_conv_term *terms, long cap, int *ncol_out);

/* x [T][nf], W [nfilter][winlen][nf], b [nfilter] -> out [ceil(T/stride)][nfilter]. */
int ffo_convolution(const float *x, int T, int nf, const float *W, const float *b, int nfilter,
                    int winlen, int stride, int act, float *out);

/* X [T][K], W [N][K], b [N] -> out [T][N]   (reference affine_map). */
void ffo_affine(const float *X, int T, int K, const float *W, const float *b, int N, float *out);

/* Xin [T][3S], sW [3S][S] -> h [T][S]; gate order (z,r,n)  (reference grumod_*). */
void ffo_grumod(const float *Xin, int T, int S, const float *sW, int backward, float *h);

/* Xin [T][4S], sW [4S][S] -> out [T][S]; gate order (i,f,g,o)  (reference lstm_*). */
void ffo_lstm(const float *Xin, int T, int S, const float *sW, int backward, float *out);

int ffo_nbase_from_nparam(int nparam);

/* h [T][S], W [nr][S], b [nr] -> trans [T][nr]  (reference globalnorm_flipflop).
 * If logZ_out != NULL receives the double log partition function. */
void ffo_globalnorm_flipflop(const float *h, int T, int S, const float *W, const float *b, int nr,
                             float temperature, float *trans, double *logZ_out);

/* trans [T][nr] -> path[T+1], qpath[T+1]; returns score  (reference decode_crf_flipflop). */
float ffo_decode_crf_flipflop(const float *trans, int T, int nr, int *path, float *qpath);

/* trans [T][nr] -> tpost [T][nr]  (reference transpost_crf_flipflop). */
int ffo_transpost_crf_flipflop(const float *trans, int T, int nr, int return_log, float *tpost);

/* tpost (probabilities, NOT log) [T][nr] -> trace [T+1][nstate]  (reference trace_from_posterior). */
void ffo_trace_from_posterior(const float *tpost, int T, int nr, int32_t *trace);

/* path[0..nblock) -> positions where the state changes  (reference change_positions). */
int ffo_change_positions(const int *path, int npos, int *chpos);

char ffo_phredf(float p);

/* Base/quality emission of calculate_post (reference src/flappie.c:284-292).
 * Returns number of bases; basecall/quality get a trailing NUL. */
int ffo_emit_bases(const int *path, const float *qpath, int nblock, int nbase, char *basecall,
                   char *quality);

/* Dense model bundle (field order follows guppy_model / guppy_stride5_model,
 * reference src/networks.c:150-215). */
typedef struct {
    int kind;            /* FFO_GRU / FFO_LSTM */
    int nconv;           /* 1 (GRU topology) or 3 (LSTM topology) */
    int conv_nf[3];      /* input features */
    int conv_nfilter[3];
    int conv_winlen[3];
    int conv_stride[3];
    const float *conv_W[3];  /* [nfilter][winlen][nf] */
    const float *conv_b[3];
    int size;            /* S */
    const float *iW[5];  /* [G*S][in]  (in = conv filters for layer 0, else S) */
    const float *sW[5];  /* [G*S][S] */
    const float *b[5];   /* [G*S] */
    int nparam;          /* rows of FF_W: nstate*(nbase+1) */
    const float *FF_W;   /* [nparam][S] */
    const float *FF_b;
} ffo_model;

/* Blocks produced for a signal of n samples (or -1 if the read is too short). */
int ffo_nblock(const ffo_model *m, int nsample);

/* Network forward: signal[n] -> trans [nblock][nparam].  Optional `layers_out[5]`
 * receive each recurrent layer's output [nblock][S], `conv_out` the last conv
 * output [nblock][nfilter].  Returns nblock or -1. */
int ffo_transitions(const ffo_model *m, const float *signal, int n, float temperature, float *trans,
                    float *conv_out, float **layers_out);

/* Whole path for one normalised read (reference calculate_post from flappie.c:262).
 * Returns number of bases or -1.  path/qpath: nblock+1 entries (may be NULL). */
int ffo_basecall(const ffo_model *m, const float *signal, int n, float temperature, int viterbi_only,
                 char *basecall, char *quality, float *score, int *path, float *qpath,
                 int32_t *trace);

#ifdef __cplusplus
}
#endif
/* ---- run-length ("runnie") head: layers.c:1235-1358, decode.c:893-1159, runnie.c:279-310 ---- */
double ffo_runlength_partition(const float *C, int T, int nr);
void ffo_globalnorm_runlength(const float *h, int T, int S, const float *W, const float *b, int nr,
                              float temperature, float *C, double *logZ_out);
float ffo_decode_crf_runlength(const float *param, int T, int nr, int *path);      /* path: T entries */
int ffo_transpost_crf_runlength(const float *param, int T, int nr, float *post);
int ffo_emit_runs(const int *path, const float *post, int T, int nr, char *bases, float *shape, float *scale, int *dwell);

#endif
