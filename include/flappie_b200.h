/* flappie_b200.h -- C ABI of libflappie_b200.so (B200 / sm_100a flip-flop basecalling hot path)
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  Two groups of entry points:
 *
 *  (1) DROP-INS with the reference's exact names, signatures, ownership and error
 *      behaviour, so that reference src/flappie.c links against this library unchanged
 *      (calculate_post, src/flappie.c:245-316, is their only caller):
 *        calculate_transitions        <- reference src/networks.h:36   (src/networks.c:108-111)
 *        transpost_crf_flipflop       <- reference src/decode.h:35     (src/decode.c:377-497)
 *        decode_crf_flipflop          <- reference src/decode.h:25     (src/decode.c:119-204)
 *        trace_from_posterior         <- reference src/decode.h:38     (src/decode.c:499-543)
 *        exp_activation_inplace       <- reference src/layers.h:17     (src/layers.c:56-66)
 *        nbase_from_flipflop_nparam   <- reference src/layers.h:89     (src/layers.c:1029-1032)
 *        decode_crf_runlength / transpost_crf_runlength
 *                                     <- reference src/decode.h:27,37  (src/decode.c:901-1159), runnie
 *        get_flappie_model_type / flappie_model_string / flappie_model_description
 *                                     <- reference src/networks.h:31-33 (src/networks.c:21-83)
 *        make/free_flappie_matrix, make/free_flappie_imatrix
 *                                     <- reference src/flappie_matrix.h:39-59 (results are
 *                                        callee-allocated, caller frees; src/flappie_matrix.c:20-51,142)
 *      NULL in -> NULL / NAN out, as RETURN_NULL_IF does in the reference
 *      (src/flappie_stdlib.h:44).  All arithmetic runs on the GPU; there is NO CPU
 *      fallback: without a usable CUDA device these functions warn and return NULL/NAN.
 *
 *  (2) The BATCHED extension (`ffb_*`) used by a rewired read loop (src/flappie.c:364-385):
 *      upload a weight bundle once, then push thousands of whole reads per call.
 *
 * The structs below are layout-compatible with the reference's (`_Mat`,
 * src/flappie_matrix.h:18-24; `raw_table`, src/flappie_structures.h:16-22) but do not
 * need <immintrin.h>.
 */
#ifndef FLAPPIE_B200_H
#define FLAPPIE_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference-compatible types ------------------------------------------------- */
#ifndef FLAPPIE_MATRIX_H   /* if the reference header is already included, use its types */
typedef struct {
    size_t nr, nrq, nc, stride;   /* column-major, stride = 4*nrq >= nr floats per column */
    union { void *v; float *f; } data;
} _Mat;
typedef struct {
    size_t nr, nrq, nc, stride;
    union { void *v; int32_t *f; } data;
} _iMat;
typedef _Mat *flappie_matrix;
typedef _iMat *flappie_imatrix;
typedef _Mat const *const_flappie_matrix;
typedef _iMat const *const_flappie_imatrix;
#endif

#ifndef FLAPPIE_STRUCTURES_H
typedef struct {
    char *uuid;
    size_t n;       /* untrimmed length */
    size_t start;   /* signal is raw[start .. end), already normalised */
    size_t end;
    float *raw;
} raw_table;
#endif

#ifndef NETWORKS_H
enum model_type {   /* reference src/networks.h:18-26, same values */
    FLAPPIE_MODEL_R941_NATIVE = 0,
    FLAPPIE_MODEL_R941_RNA002,
    FLAPPIE_MODEL_R941_5mC,
    FLAPPIE_MODEL_R103_NATIVE,
    FLAPPIE_MODEL_INVALID,
    RUNNIE_MODEL_R941_NATIVE,
    RUNNIE_MODEL_INVALID
};
#endif

/* ---- (1) drop-ins ---------------------------------------------------------------- */
flappie_matrix make_flappie_matrix(size_t nr, size_t nc);
flappie_matrix free_flappie_matrix(flappie_matrix mat);
flappie_imatrix make_flappie_imatrix(size_t nr, size_t nc);
flappie_imatrix free_flappie_imatrix(flappie_imatrix mat);

enum model_type get_flappie_model_type(const char *modelstr);   /* also accepts "r10C_pcr" */
const char *flappie_model_string(const enum model_type model);
const char *flappie_model_description(const enum model_type model);

flappie_matrix calculate_transitions(const raw_table signal, float temperature, enum model_type model);
flappie_matrix transpost_crf_flipflop(const_flappie_matrix trans, bool return_log);
float decode_crf_flipflop(const_flappie_matrix trans, bool combine_stays, int *path, float *qpath);
flappie_imatrix trace_from_posterior(flappie_matrix tpost);
void exp_activation_inplace(flappie_matrix C);
size_t nbase_from_flipflop_nparam(size_t nparam);
/* the other symbols of the replaced files that the reference's remaining sources link against (flappie.c:284,
 * runnie.c:262, fast5_interface.c:138): reference src/decode.h:21, src/layers.h:95, src/flappie_matrix.h:63 */
size_t change_positions(int const *path, size_t npos, int *chpos);
size_t nbase_from_crf_runlength_nparam(size_t nparam);
int32_t *array_from_flappie_imatrix(const_flappie_imatrix mat);
/* run-length ("runnie") decoding, reference src/decode.h:27,37 (src/decode.c:901-1159); path has nblock entries */
float decode_crf_runlength(const_flappie_matrix param, int *path);
flappie_matrix transpost_crf_runlength(const_flappie_matrix param);

/* ---- (2) batched extension ------------------------------------------------------- */

#define FFB_KIND_GRU 0    /* guppy_model          (reference src/networks.c:150-177) */
#define FFB_KIND_LSTM 1   /* guppy_stride5_model  (reference src/networks.c:180-215) */
#define FFB_KIND_RUNLENGTH 2   /* the same 23-matrix bundle with the run-length ("runnie") head:
                                * runlength5_guppy_transitions, reference src/networks.c:675-722 */

#define FFB_OK 0
#define FFB_ERR_ARG -1
#define FFB_ERR_CUDA -2
#define FFB_ERR_NOMEM -3
#define FFB_ERR_UNSUPPORTED -4

typedef struct ffb_model ffb_model;   /* device-resident, pre-packed weight arena */
typedef struct ffb_ctx ffb_ctx;       /* one per (device, stream): workspaces + plans */

/* Number of usable CUDA devices (0 if none). */
int ffb_device_count(void);
const char *ffb_last_error(void);
const char *ffb_version(void);

/* Weight bundle in the reference's own field order:
 *   kind GRU : conv_W, conv_b, {iW, sW, b} x 5 (B1,F2,B3,F4,B5), FF_W, FF_b         (19 mats)
 *   kind LSTM: conv1_W, conv1_b, conv2_W, conv2_b, conv3_W, conv3_b, {iW,sW,b} x 5,
 *              FF_W, FF_b                                                           (23 mats)
 * `mats[i]` are reference `_Mat`s exactly as the generated model headers define them
 * (convolution filters: nr = nf4*winlen - nf4 + nf).  conv_stride has 1 or 3 entries. */
ffb_model *ffb_model_create(int device, int kind, const _Mat *const *mats, int nmat,
                            const int *conv_stride, int nconv);
/* The same from a weight bundle file ("FFBW1": kind, nconv, strides, nmat, then nmat x {uint64 nr, uint64 nc, padded
 * column-major floats}; written by flappie_b200.model.FlipflopModel.save_bundle).  calculate_transitions() loads
 * $FLAPPIE_B200_MODELS/<model name>.ffbw by itself (on device $FLAPPIE_B200_DEVICE, default 0) when nothing has been
 * registered for a model -- which is all the reference's own main() needs to run on this library. */
ffb_model *ffb_model_load(const char *path, int device);
void ffb_model_destroy(ffb_model *m);
int ffb_model_size(const ffb_model *m);      /* S */
int ffb_model_nparam(const ffb_model *m);    /* rows of trans: nstate*(nbase+1) */
int ffb_model_stride(const ffb_model *m);    /* product of conv strides */
/* blocks for a signal of `nsample` samples (iceil per convolution, reference
 * src/layers.c:204), or -1 if shorter than a filter window (the reference's index
 * arithmetic underflows there, src/layers.c:262). */
long ffb_model_nblock(const ffb_model *m, long nsample);

/* Bind a model to a reference enum value for the per-read drop-in calculate_transitions(). */
int ffb_register_model(enum model_type which, ffb_model *m);

/* `stream`: a cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream),
 * or NULL to let the context create its own non-blocking stream. */
ffb_ctx *ffb_create(ffb_model *m, void *stream);
void ffb_destroy(ffb_ctx *c);

#define FFB_FLAG_VITERBI_ONLY 1u   /* --viterbi: decode trans directly (reference src/flappie.c:278-283) */
#define FFB_FLAG_WANT_TRACE 2u     /* compute the u8 state trace (reference src/flappie.c:299-300) */
#define FFB_FLAG_WANT_TRANS 4u     /* copy trans (and tpost) back to the host */
#define FFB_FLAG_KEEP_LAYERS 8u    /* keep every layer's output on the device for ffb_debug_fetch() */
#define FFB_FLAG_FP32_SIMT 16u     /* force the fp32 CUDA-core GEMM / recurrence kernels */
#define FFB_FLAG_FP32_CONV 32u     /* keep every convolution on the fp32 CUDA-core kernels: the tensor-core last convolution
                                    * of the LSTM topology carries 22-bit operands, fine for med-MAD normalised signal
                                    * (activations O(1)) but 1.6e-4 on trans for --delta input (activations O(100));
                                    * ffb_upload_raw sets it by itself when delta != 0 */
#define FFB_FLAG_REVERSE 64u       /* --reverse: bases / quals of every read come out reversed (reference src/flappie.c:293-297) */

/* One batch of whole reads.  All pointers are HOST memory owned by the caller.
 * signal      : concatenated normalised samples of all reads
 * sig_off[n]  : start of read n in `signal` (n_reads+1 entries)
 * Outputs (any may be NULL):
 * blk_off     : n_reads+1 entries; read n owns blocks [blk_off[n], blk_off[n+1])
 * path, qpath : sum(T_n + 1) entries; read n starts at blk_off[n] + n
 * score       : n_reads   (NAN for a rejected read)
 * trans,tpost : sum(T_n) * nparam floats, [block][nparam] row-major (= reference columns)
 * trace       : sum(T_n + 1) * nstate bytes, read n starts at (blk_off[n] + n) * nstate
 * bases, quals, nbases (all three or none): the called bases and their quality characters, emitted on the device
 *               (reference src/decode.c:66-79, src/flappie.c:284-297, src/util.h:285-305).  bases / quals: sum(T_n + 1)
 *               chars, read n's NUL-terminated string starts at blk_off[n] + n; nbases[n] = its length.  Identical to
 *               ffb_emit_bases() on path / qpath, which then need not be copied back at all.
 * Returns FFB_OK or a negative error; a read shorter than a filter window gets T_n = 0. */
typedef struct {
    const float *signal;
    const int64_t *sig_off;
    int64_t n_reads;
    float temperature;
    uint32_t flags;
    int64_t *blk_off;
    int32_t *path;
    float *qpath;
    float *score;
    float *trans;
    float *tpost;
    uint8_t *trace;
    float *rle_params;   /* FFB_KIND_RUNLENGTH models only (may be NULL): sum(T_n) * 8 floats, [block][shape ACGT, scale ACGT];
                          * for those models path[] holds the run-length states (T_n entries + one -1), qpath zeros */
    char *bases;
    char *quals;
    int32_t *nbases;
} ffb_batch;

/* Upload, run the whole hot path, download, synchronise. */
int ffb_basecall_batch(ffb_ctx *c, const ffb_batch *b);

/* Split version for overlap / device-resident timing:
 *   ffb_upload  : H2D of signal (async on the context stream) + host-side planning
 *   ffb_forward : every kernel of the path, async, inputs and outputs stay in HBM
 *   ffb_download: D2H of the requested outputs + stream synchronise */
int ffb_upload(ffb_ctx *c, const ffb_batch *b);
int ffb_forward(ffb_ctx *c);
int ffb_download(ffb_ctx *c, const ffb_batch *b);
int ffb_sync(ffb_ctx *c);

/* Optional: size the context's (grow-only) device workspaces for batches of n_reads reads of samples_per_read samples
 * ahead of the first batch, e.g. at start-up next to the weight upload.  `flags` as in ffb_batch.flags. */
int ffb_reserve(ffb_ctx *c, int64_t n_reads, int64_t samples_per_read, uint32_t flags);

/* Raw reads: the signal preparation of calculate_post (reference src/flappie.c:251-259) on the device --
 * trim_and_segment_raw (src/flappie_common.c:13-81, chunk MADs + threshold quantile), then
 * medmad_normalise_array (src/util.c:198-212), or difference_array + shift_scale_array when delta != 0
 * (src/util.c:215-223,278-287) -- followed by the same plan as ffb_upload.
 * raw         : concatenated raw samples (pA, as read_raw returns them), HOST memory
 * raw_off[n]  : start of read n in `raw` (n_reads+1 entries)
 * trim_start, trim_end, varseg_chunk, varseg_thresh, delta: the reference CLI options
 *               (--trim 200:10, --segmentation 100:0.0, --delta 0.0; src/flappie.c:100-110)
 * start, end  : outputs (may be NULL), n_reads entries: the kept range of each read; start >= end means the
 *               reference would have dropped the read -- it gets T_n = 0 and score NAN.
 * `b` supplies n_reads (must match), temperature, flags and the output pointers; its signal / sig_off are
 * ignored.  path/qpath/trace must be sized for the UNTRIMMED lengths (an upper bound on the block count). */
typedef struct {
    const float *raw;
    const int64_t *raw_off;
    int64_t n_reads;
    int64_t trim_start, trim_end, varseg_chunk;
    float varseg_thresh;
    float delta;
    int64_t *start;
    int64_t *end;
} ffb_raw_batch;
int ffb_upload_raw(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b);
int ffb_basecall_raw_batch(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b);

/* Pipelined use (the `ffgpu_submit` / `ffgpu_collect` pair of SURVEY section 8b): submit = upload (or raw upload),
 * every kernel and the D2H copies enqueued on the context's stream, returning WITHOUT waiting for the results;
 * ffb_collect waits for that context and finalises its outputs.  With two contexts on two streams the host reads,
 * plans and uploads batch i+1 while the device works on batch i.  Output buffers should be pinned host memory
 * (cudaHostAlloc / torch pin_memory) for the copies to be asynchronous; they belong to the library until
 * ffb_collect returns. */
int ffb_submit_batch(ffb_ctx *c, const ffb_batch *b);
int ffb_submit_raw_batch(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b);
/* ffb_submit_raw_batch in two halves.  The plan of a raw batch needs the trimmed lengths back from the device: `begin`
 * enqueues the raw upload, the trimming kernels and the copy back of their bounds and returns at once; `finish` waits for
 * those bounds, plans, and enqueues everything else (it is what blocks in ffb_submit_raw_batch).  A host thread can read the
 * next batch's files in between.  `rb`, `b` and what they point to belong to the library until ffb_collect returns. */
int ffb_submit_raw_begin(ffb_ctx *c, const ffb_raw_batch *rb, const ffb_batch *b);
int ffb_submit_raw_finish(ffb_ctx *c);
int ffb_collect(ffb_ctx *c, const ffb_batch *b);
/* zero-filled page-locked host memory (NULL on failure), for callers that do not link the CUDA runtime */
void *ffb_alloc_pinned(size_t bytes);
void ffb_free_pinned(void *p);

/* Introspection for tests / bench. */
int64_t ffb_total_blocks(const ffb_ctx *c);
int64_t ffb_launch_count(const ffb_ctx *c);         /* kernels launched by this context so far */
/* Time the kernels of one ffb_forward() by group with CUDA events on the context stream.
 * ms[0]=conv ms[1]=input GEMMs ms[2]=recurrent ms[3]=output layer+logZ ms[4]=decode; returns FFB_OK. */
int ffb_forward_timed(ffb_ctx *c, float ms[8]);
/* what: 0 = last conv output [Ttot][S]; 1..5 = recurrent layer output [Ttot][S] (needs
 * FFB_FLAG_KEEP_LAYERS); 6 = trans [Ttot][nparam]; 7 = logZ (double, n_reads); 8 = the normalised signal
 * the network reads (concatenated kept ranges).  Copies up to
 * `bytes` to host `dst`; returns bytes copied or negative error. */
int64_t ffb_debug_fetch(ffb_ctx *c, int what, void *dst, int64_t bytes);
/* Libraries built with -DFFB_RNN_PROFILE only (FFB_ERR_UNSUPPORTED otherwise): CUDA-event times of the LAST ordinary
 * ffb_forward on this context, in the schedule as it runs -- ms[0] convolutions, ms[1] first input GEMM .. last recurrent
 * layer, ms[2] output layer, ms[3] decoding (tools/step_timeline.py). */
int ffb_debug_group_times(ffb_ctx *c, float ms[4]);

/* The host-side schedule of the tensor recurrent kernel, callable without a device (tests, capacity planning): reads
 * sorted by length form groups of 16; each of the n_clusters x slots slots gets a list of groups, longest first to the
 * least-loaded slot.  T[n] = blocks of read n.  order: 16 * groups entries (-1 padded), slot_off: n_clusters * slots + 1,
 * slot_list: groups entries; any of them may be NULL.  Returns the number of groups, or -1. */
int64_t ffb_plan_schedule(const int64_t *T, int64_t n_reads, int max_clusters, int slots_max, int can_stream,
                          int32_t *order, int32_t *slot_off, int32_t *slot_list, int *n_clusters, int *slots);

/* Base / quality emission of calculate_post (reference src/flappie.c:284-297,
 * src/decode.c:66-79, src/util.h:285-305) on the host: returns the number of bases
 * written to basecall/quality (each needs nblock+1 chars, NUL-terminated). */
int ffb_emit_bases(const int32_t *path, const float *qpath, int64_t nblock, int nbase, bool reverse,
                   char *basecall, char *quality);
/* The quality character is a step function of qpath: out[k] (ascending) = the smallest qpath value whose character is
 * >= 34 + k under the host's libm; the device emission compares against this table.  Returns the table length. */
int ffb_phred_table(float *out, int cap);
/* The run loop of runnie's calculate_post (reference src/runnie.c:279-310): returns the number of runs written to
 * bases (NUL-terminated) / shape / scale / dwell (each needs nblock + 1 entries). */
int64_t ffb_emit_runs(const int32_t *path, const float *rle_params, int64_t nblock, int nbase, char *bases,
                      float *shape, float *scale, int32_t *dwell);

#ifdef __cplusplus
}
#endif
#endif /* FLAPPIE_B200_H */
