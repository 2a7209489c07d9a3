/* flappie_main.c -- the `flappie` command line on the B200 path (host code in C, reference src/flappie.c).
 *
 * Same options, defaults, model names and record formats as the reference binary; the per-file calculate_post
 * loop (src/flappie.c:364-385) is rewired to: read up to --batch raw reads -> ffb_basecall_raw_batch (trimming,
 * normalisation, network, decoding on the device) -> emit bases -> print in input order.
 *
 * Differences forced by this image: no libhdf5, so reads come from <name>.f32 / <name>.crp files instead of
 * fast5 (ffb_host.h) and --trace is refused; weights come from a bundle file (--weights, or
 * $FLAPPIE_B200_MODELS/<model>.ffbw) because the reference's compiled-in .mdl headers are git-LFS objects.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <getopt.h>
#include <glob.h>
#include <libgen.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "ffb_host.h"

/* The same driver built with -DFFB_RUNNIE is `runnie` (reference src/runnie.c): run-length model, `.run` text output
 * ("# uuid" then one "base<TAB>shape<TAB>scale<TAB>run" line per run, src/runnie.c:277-310). */
#ifdef FFB_RUNNIE
#define DEFAULT_MODEL RUNNIE_MODEL_R941_NATIVE
#define DEFAULT_MODEL_NAME "rle_r941_native"
#define PROGRAM "runnie"
#else
#define DEFAULT_MODEL FLAPPIE_MODEL_R941_NATIVE
#define DEFAULT_MODEL_NAME "r941_native"
#define PROGRAM "flappie"
#endif

struct arguments {
    float delta;
    char *trace;
    enum ffb_outformat outformat;
    int limit;
    enum model_type model;
    const char *model_name;
    FILE *output;
    const char *prefix;
    bool reverse;
    float temperature;
    int trim_start, trim_end;
    int varseg_chunk;
    float varseg_thresh;
    bool viterbi_only;
    bool uuid;
    /* additions */
    const char *weights;
    int batch;
    int device;
};

/* defaults of the reference, src/flappie.c:93-112 */
static struct arguments args = {
    .delta = 0.0f, .trace = NULL, .outformat = FFB_OUT_FASTQ, .limit = 0, .model = DEFAULT_MODEL,
    .model_name = DEFAULT_MODEL_NAME, .output = NULL, .prefix = "", .reverse = false, .temperature = 1.0f,
    .trim_start = 200, .trim_end = 10, .varseg_chunk = 100, .varseg_thresh = 0.0f, .viterbi_only = false,
    .uuid = true, .weights = NULL, .batch = 1024, .device = 0};

static void die(const char *fmt, const char *arg) {
    fprintf(stderr, PROGRAM ": ");
    fprintf(stderr, fmt, arg);
    fputc('\n', stderr);
    exit(EXIT_FAILURE);
}

static void print_models(FILE *fh) {
#ifdef FFB_RUNNIE
    fprintf(fh, "%10s : %s  %s\n", flappie_model_string(RUNNIE_MODEL_R941_NATIVE), flappie_model_description(RUNNIE_MODEL_R941_NATIVE), "(default)");
    return;
#endif
    for (int mdl = 0; mdl < (int)FLAPPIE_MODEL_INVALID; mdl++)
        fprintf(fh, "%10s : %s  %s\n", flappie_model_string((enum model_type)mdl), flappie_model_description((enum model_type)mdl),
                (DEFAULT_MODEL == mdl) ? "(default)" : "");
}

static void usage(FILE *fh) {
    fputs("Usage: flappie [OPTION...] signal [signal ...]\n"
          "Flappie basecaller -- basecall from raw signal (B200 build)\n\n"
          "  -d, --delta=factor         Using delta samples model with scaling factor\n"
          "  -f, --format=format        Format to output reads (fasta, fastq or sam)\n"
          "  -l, --limit=nreads         Maximum number of reads to call (0 is unlimited)\n"
          "  -m, --model=name           Model to use (\"help\" to list)\n"
          "  -o, --output=filename      Write to file rather than stdout\n"
          "  -p, --prefix=string        Prefix to append to name of each read\n"
          "  -r, --reverse, --no-reverse  Reverse output base calls\n"
          "      --temperature=factor   Temperature for weights\n"
          "  -t, --trim=start:end       Number of samples to trim, as start:end\n"
          "  -T, --trace=filename       Dump trace to HDF5 file (needs libhdf5: refused by this build)\n"
          "      --segmentation=chunk:percentile  Chunk size and percentile for variance based segmentation\n"
          "  -v, --viterbi, --no-viterbi, --fb  Use viterbi decoding only / forward-backward followed by viterbi\n"
          "      --uuid, --no-uuid      Output UUID / read file name\n"
          "      --weights=file         Weight bundle (default $FLAPPIE_B200_MODELS/<model>.ffbw)\n"
          "      --batch=nreads         Reads per device batch (default 1024)\n"
          "      --device=index         CUDA device (default 0)\n", fh);
}

enum { OPT_SEG = 3, OPT_NOREV = 6, OPT_TEMP, OPT_NOVIT, OPT_FB, OPT_LIC, OPT_LIC2, OPT_H5C, OPT_H5K, OPT_UUID, OPT_NOUUID,
       OPT_WEIGHTS = 1000, OPT_BATCH, OPT_DEVICE, OPT_HELP };

static const struct option long_opts[] = {
    {"delta", required_argument, 0, 'd'}, {"format", required_argument, 0, 'f'}, {"limit", required_argument, 0, 'l'},
    {"model", required_argument, 0, 'm'}, {"output", required_argument, 0, 'o'}, {"prefix", required_argument, 0, 'p'},
    {"reverse", no_argument, 0, 'r'}, {"no-reverse", no_argument, 0, OPT_NOREV}, {"temperature", required_argument, 0, OPT_TEMP},
    {"trim", required_argument, 0, 't'}, {"trace", required_argument, 0, 'T'}, {"licence", no_argument, 0, OPT_LIC},
    {"license", no_argument, 0, OPT_LIC2}, {"segmentation", required_argument, 0, OPT_SEG}, {"viterbi", no_argument, 0, 'v'},
    {"no-viterbi", no_argument, 0, OPT_NOVIT}, {"fb", no_argument, 0, OPT_FB}, {"hdf5-compression", required_argument, 0, OPT_H5C},
    {"hdf5-chunk", required_argument, 0, OPT_H5K}, {"uuid", no_argument, 0, OPT_UUID}, {"no-uuid", no_argument, 0, OPT_NOUUID},
    {"weights", required_argument, 0, OPT_WEIGHTS}, {"batch", required_argument, 0, OPT_BATCH},
    {"device", required_argument, 0, OPT_DEVICE}, {"help", no_argument, 0, OPT_HELP}, {0, 0, 0, 0}};

static void parse_args(int argc, char **argv) {
    int key;
    char *tok;
    while ((key = getopt_long(argc, argv, "d:f:l:m:o:p:rt:T:v", long_opts, NULL)) != -1) {
        switch (key) {
        case 'd': args.delta = atof(optarg); break;
        case 'f':
            args.outformat = ffb_get_outformat(optarg);
            if (FFB_OUT_INVALID == args.outformat) die("Unrecognised output format \"%s\".", optarg);
            break;
        case 'l': args.limit = atoi(optarg); break;
        case 'm':
            if (0 == strcasecmp(optarg, "help")) { print_models(stdout); exit(EXIT_SUCCESS); }
            args.model = get_flappie_model_type(optarg);
            args.model_name = optarg;
#ifdef FFB_RUNNIE
            if (RUNNIE_MODEL_R941_NATIVE != args.model) {
#else
            if (FLAPPIE_MODEL_INVALID == args.model || args.model >= FLAPPIE_MODEL_INVALID) {
#endif
                fprintf(stdout, "Invalid Flappie model \"%s\".\n", optarg);
                print_models(stdout);
                exit(EXIT_FAILURE);
            }
            break;
        case 'o':
            args.output = fopen(optarg, "w");
            if (!args.output) die("Failed to open \"%s\" for output.", optarg);
            break;
        case 'p': args.prefix = optarg; break;
        case 'r': args.reverse = true; break;
        case 't':
            args.trim_start = atoi(strtok(optarg, ":"));
            tok = strtok(NULL, ":");
            args.trim_end = tok ? atoi(tok) : args.trim_start;
            if (args.trim_start < 0 || args.trim_end < 0) die("--trim must not be negative%s", "");
            break;
        case 'T': args.trace = optarg; break;
        case 'v': args.viterbi_only = true; break;
        case OPT_SEG:
            args.varseg_chunk = atoi(strtok(optarg, ":"));
            tok = strtok(NULL, ":");
            if (!tok) die("--segmentation should be of form chunk:percentile%s", "");
            args.varseg_thresh = atof(tok) / 100.0;
            break;
        case OPT_NOREV: args.reverse = false; break;
        case OPT_TEMP:
            args.temperature = atof(optarg);
            if (!isfinite(args.temperature) || args.temperature <= 0.0f) die("--temperature must be positive%s", "");
            break;
        case OPT_NOVIT: case OPT_FB: args.viterbi_only = false; break;
        case OPT_LIC: case OPT_LIC2:
            fputs("flappie_b200: see the reference's LICENCE.txt for the Oxford Nanopore Technologies Public License\n", stdout);
            exit(EXIT_SUCCESS);
        case OPT_H5C: case OPT_H5K: break;   /* only meaningful with --trace */
        case OPT_UUID: args.uuid = true; break;
        case OPT_NOUUID: args.uuid = false; break;
        case OPT_WEIGHTS: args.weights = optarg; break;
        case OPT_BATCH: args.batch = atoi(optarg) > 0 ? atoi(optarg) : 1; break;
        case OPT_DEVICE: args.device = atoi(optarg); break;
        case OPT_HELP: usage(stdout); exit(EXIT_SUCCESS);
        default: usage(stderr); exit(EXIT_FAILURE);
        }
    }
    if (optind >= argc) { usage(stderr); exit(EXIT_FAILURE); }
}

/* ---- a batch of raw reads waiting for the device ---- */
struct pending {
    char **name;            /* basename of each file (printed as "filename", and as the id with --no-uuid) */
    char **uuid;
    float *raw;             /* concatenated samples */
    int64_t *raw_off;
    int n, cap;
    size_t raw_cap;
};

static void pending_add(struct pending *p, const char *path, float *sig, long n) {
    if (p->n == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 256;
        p->name = realloc(p->name, sizeof(char *) * (size_t)p->cap);
        p->uuid = realloc(p->uuid, sizeof(char *) * (size_t)p->cap);
        p->raw_off = realloc(p->raw_off, sizeof(int64_t) * ((size_t)p->cap + 1));
        if (!p->name || !p->uuid || !p->raw_off) die("out of memory%s", "");
    }
    if (p->n == 0) p->raw_off[0] = 0;
    const size_t need = (size_t)p->raw_off[p->n] + (size_t)n;
    if (need > p->raw_cap) {
        p->raw_cap = need * 2 + 4096;
        p->raw = realloc(p->raw, sizeof(float) * p->raw_cap);
        if (!p->raw) die("out of memory%s", "");
    }
    memcpy(p->raw + p->raw_off[p->n], sig, sizeof(float) * (size_t)n);
    char *tmp = strdup(path);
    p->name[p->n] = strdup(basename(tmp));
    free(tmp);
    /* fast5 files carry a read uuid attribute; a bare signal file has none: its stem stands in */
    p->uuid[p->n] = strdup(p->name[p->n]);
    char *dot = strrchr(p->uuid[p->n], '.');
    if (dot) *dot = 0;
    p->raw_off[p->n + 1] = (int64_t)need;
    p->n++;
}

static void pending_clear(struct pending *p) {
    for (int i = 0; i < p->n; i++) { free(p->name[i]); free(p->uuid[i]); }
    p->n = 0;
}

/* One batch in flight on one context: calculate_post for every pending read at once (submitted without waiting),
 * then -- at collect time -- the reference's per-read printing in input order.  main() keeps two of these going, so
 * the files of batch i+1 are read while the device works on batch i. */
struct inflight {
    ffb_ctx *ctx;
    struct pending pend;
    bool busy;
    int64_t *blk_off, *start, *end;
    int32_t *path;
    float *qpath, *score, *rle;
    ffb_batch b;
};

static void submit_batch(struct inflight *f, ffb_model *model) {
    struct pending *p = &f->pend;
    if (p->n == 0) return;
    const int n = p->n;
    int64_t tot_blocks = 0;
    for (int i = 0; i < n; i++) {
        const long t = ffb_model_nblock(model, (long)(p->raw_off[i + 1] - p->raw_off[i]));
        tot_blocks += t > 0 ? t : 0;                  /* upper bound: the kept range is shorter */
    }
    f->blk_off = calloc((size_t)n + 1, sizeof(int64_t));
    f->start = calloc((size_t)n, sizeof(int64_t));
    f->end = calloc((size_t)n, sizeof(int64_t));
    /* page-locked, so that the device-to-host copies really are asynchronous */
    f->path = ffb_alloc_pinned((size_t)(tot_blocks + n) * sizeof(int32_t));
    f->qpath = ffb_alloc_pinned((size_t)(tot_blocks + n) * sizeof(float));
    f->score = ffb_alloc_pinned((size_t)n * sizeof(float));
#ifdef FFB_RUNNIE
    f->rle = ffb_alloc_pinned((size_t)(tot_blocks + 1) * 8 * sizeof(float));
    if (!f->rle) die("out of memory%s", "");
#endif
    if (!f->blk_off || !f->start || !f->end || !f->path || !f->qpath || !f->score) die("out of memory%s", "");
    ffb_raw_batch rb = {.raw = p->raw, .raw_off = p->raw_off, .n_reads = n, .trim_start = args.trim_start, .trim_end = args.trim_end,
                        .varseg_chunk = args.varseg_chunk, .varseg_thresh = args.varseg_thresh, .delta = args.delta,
                        .start = f->start, .end = f->end};
    memset(&f->b, 0, sizeof f->b);
    f->b.n_reads = n; f->b.temperature = args.temperature;
    f->b.flags = args.viterbi_only ? FFB_FLAG_VITERBI_ONLY : 0;
    f->b.blk_off = f->blk_off; f->b.path = f->path; f->b.qpath = f->qpath; f->b.score = f->score;
    f->b.rle_params = f->rle;
    if (ffb_submit_raw_batch(f->ctx, &rb, &f->b) != FFB_OK) die("device basecall failed: %s", ffb_last_error());
    f->busy = true;
}

static void collect_batch(struct inflight *f, ffb_model *model) {
    if (!f->busy) return;
    struct pending *p = &f->pend;
    const int n = p->n;
    if (ffb_collect(f->ctx, &f->b) != FFB_OK) die("device basecall failed: %s", ffb_last_error());
    const int nbase = (int)nbase_from_flipflop_nparam((size_t)ffb_model_nparam(model));
    for (int i = 0; i < n; i++) {
        const int64_t nblock = f->blk_off[i + 1] - f->blk_off[i];
        if (nblock <= 0) {
            fprintf(stderr, PROGRAM ": No basecall returned for %s\n", p->name[i]);   /* src/flappie.c:370-373 */
            continue;
        }
#ifdef FFB_RUNNIE
        {
            char *bases = calloc((size_t)nblock + 2, 1);
            float *shape = calloc((size_t)nblock + 1, sizeof(float)), *scale = calloc((size_t)nblock + 1, sizeof(float));
            int32_t *dwell = calloc((size_t)nblock + 1, sizeof(int32_t));
            if (!bases || !shape || !scale || !dwell) die("out of memory%s", "");
            const int64_t nrun = ffb_emit_runs(f->path + f->blk_off[i] + i, f->rle + f->blk_off[i] * 8, nblock, nbase, bases, shape, scale, dwell);
            fprintf(args.output, "# %s\n", p->uuid[i]);                                  /* src/runnie.c:277 */
            for (int64_t r = 0; r < nrun; r++) fprintf(args.output, "%c\t%f\t%f\t%d\n", bases[r], shape[r], scale[r], dwell[r]);
            free(bases); free(shape); free(scale); free(dwell);
            continue;
        }
#endif
        char *basecall = calloc((size_t)nblock + 2, 1), *quality = calloc((size_t)nblock + 2, 1);
        if (!basecall || !quality) die("out of memory%s", "");
        const int nb = ffb_emit_bases(f->path + f->blk_off[i] + i, f->qpath + f->blk_off[i] + i, nblock, nbase, args.reverse, basecall, quality);
        ffb_read_result res = {.score = f->score[i], .n = (size_t)(p->raw_off[i + 1] - p->raw_off[i]), .start = (size_t)f->start[i],
                               .end = (size_t)f->end[i], .basecall = basecall, .quality = quality, .basecall_length = (size_t)(nb > 0 ? nb : 0),
                               .nblock = (size_t)nblock};
        ffb_fprintf_read(args.outformat, args.output, p->uuid[i], p->name[i], args.uuid, args.prefix, &res);
        free(basecall); free(quality);
    }
    free(f->blk_off); free(f->start); free(f->end);
    ffb_free_pinned(f->path); ffb_free_pinned(f->qpath); ffb_free_pinned(f->score); ffb_free_pinned(f->rle);
    f->rle = NULL;
    pending_clear(p);
    f->busy = false;
}

int main(int argc, char **argv) {
    parse_args(argc, argv);
    if (!args.output) args.output = stdout;
    if (args.trace) die("--trace %s: HDF5 output needs libhdf5, which this build does not have", args.trace);
    if (ffb_device_count() <= args.device) die("no CUDA device: flappie_b200 has no CPU fallback%s", "");

    char wpath[4096];
    if (!args.weights) {
        const char *dir = getenv("FLAPPIE_B200_MODELS");
        if (!dir) die("no weights: give --weights <bundle> or set FLAPPIE_B200_MODELS (model %s)", args.model_name);
        snprintf(wpath, sizeof wpath, "%s/%s.ffbw", dir, args.model_name);
        args.weights = wpath;
    }
    ffb_bundle bundle;
    if (ffb_bundle_load(args.weights, &bundle) != 0) die("cannot read weight bundle \"%s\"", args.weights);
    ffb_model *model = ffb_bundle_to_model(&bundle, args.device);
    if (!model) die("weight bundle rejected: %s", ffb_last_error());
    ffb_bundle_free(&bundle);
    struct inflight fl[2];
    memset(fl, 0, sizeof fl);
    for (int k = 0; k < 2; k++) {
        fl[k].ctx = ffb_create(model, NULL);
        if (!fl[k].ctx) die("ffb_create: %s", ffb_last_error());
    }
    int cur = 0;                                    /* the batch being filled; the other one may be on the device */
    int reads_started = 0;
    for (int fn = optind; fn < argc; fn++) {
        if (args.limit > 0 && reads_started >= args.limit) continue;
        /* files and directories, through the system glob as the reference does (src/flappie.c:338-362) */
        glob_t globbuf;
        char *pattern = calloc(strlen(argv[fn]) + 16, 1);
        strcpy(pattern, argv[fn]);
        DIR *dirp = opendir(argv[fn]);
        const bool is_dir = dirp != NULL;
        if (is_dir) { strcat(pattern, "/*"); closedir(dirp); }   /* a directory: every signal file in it, sorted by name */
        const int globret = glob(pattern, 0, NULL, &globbuf);
        free(pattern);
        if (0 != globret) {
            if (GLOB_NOMATCH == globret) fprintf(stderr, PROGRAM ": File or directory \"%s\" does not exist or no signal files found.\n", argv[fn]);
            globfree(&globbuf);
            continue;
        }
        for (size_t k = 0; k < globbuf.gl_pathc; k++) {
            const char *filename = globbuf.gl_pathv[k];
            if (is_dir && !ffb_is_signal_file(filename)) continue;
            if (args.limit > 0 && reads_started >= args.limit) continue;
            reads_started += 1;
            float *sig = NULL;
            const long n = ffb_read_raw_file(filename, &sig);
            if (n == -2) { fprintf(stderr, PROGRAM ": %s: fast5 input needs libhdf5, which this build does not have\n", filename); continue; }
            if (n <= 0) { fprintf(stderr, PROGRAM ": No basecall returned for %s\n", filename); free(sig); continue; }
            pending_add(&fl[cur].pend, filename, sig, n);
            free(sig);
            if (fl[cur].pend.n >= args.batch) {
                submit_batch(&fl[cur], model);
                cur ^= 1;
                collect_batch(&fl[cur], model);     /* the older batch: print it, then refill its slot */
            }
        }
        globfree(&globbuf);
    }
    collect_batch(&fl[cur ^ 1], model);             /* older batch first: records stay in input order */
    submit_batch(&fl[cur], model);
    collect_batch(&fl[cur], model);

    ffb_destroy(fl[0].ctx);
    ffb_destroy(fl[1].ctx);
    ffb_model_destroy(model);
    if (stdout != args.output) fclose(args.output);
    return EXIT_SUCCESS;
}
