/* flappie_main.c -- the `flappie` command line on the B200 path (host code in C, reference src/flappie.c).
 *
 * Same options, defaults, model names and record formats as the reference binary.  The reference's parallelism is "run N
 * processes" (README.md:80-83, an OpenMP loop over files in src/flappie.c:364-385); here ONE process drives any number of
 * GPUs of the box (--devices 0-7).  The per-file calculate_post loop is rewired to:
 *
 *   window  = the next (devices x --batch) files, in input order
 *   read    : every device thread reads its share of the window's files                      (parallel host I/O)
 *   deal    : reads sorted by length, longest first to the least-loaded device (LPT)         (main thread)
 *   submit  : each device thread packs its batch into pinned memory and enqueues the raw upload + trimming
 *             (ffb_submit_raw_begin), reads its share of the NEXT window while the device trims, then plans and enqueues
 *             normalisation + network + decoding + base/quality emission + the device-to-host copies
 *             (ffb_submit_raw_finish) -- two contexts per device, so the GPUs work on window w-1 meanwhile
 *   collect : wait for the older window's batch and format its records into memory           (device threads)
 *   print   : the records of a window in INPUT ORDER, whatever device called them             (main thread, one window
 *             behind the formatting, concurrent with the device threads)
 *
 * Reads never cross devices and there is no collective: basecalling shards embarrassingly.
 *
 * Differences forced by this image: no libhdf5, so reads come from <name>.f32 / <name>.crp files instead of fast5
 * (ffb_host.h) and --trace writes a flat binary file instead of HDF5 (format in ffb_host.h); weights come from a bundle
 * file (--weights, or $FLAPPIE_B200_MODELS/<model>.ffbw) because the reference's compiled-in .mdl headers are git-LFS
 * objects.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <getopt.h>
#include <glob.h>
#include <libgen.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

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
#define MAX_DEVICES 64

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
    int ndev;
    int devices[MAX_DEVICES];
    bool stats;
};

/* defaults of the reference, src/flappie.c:93-112 */
static struct arguments args = {
    .delta = 0.0f, .trace = NULL, .outformat = FFB_OUT_FASTQ, .limit = 0, .model = DEFAULT_MODEL,
    .model_name = DEFAULT_MODEL_NAME, .output = NULL, .prefix = "", .reverse = false, .temperature = 1.0f,
    .trim_start = 200, .trim_end = 10, .varseg_chunk = 100, .varseg_thresh = 0.0f, .viterbi_only = false,
    .uuid = true, .weights = NULL, .batch = 1024, .ndev = 1, .devices = {0}, .stats = false};

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
          "  -T, --trace=filename       Dump trace to file (flat binary records, see ffb_host.h: this build has no libhdf5)\n"
          "      --segmentation=chunk:percentile  Chunk size and percentile for variance based segmentation\n"
          "  -v, --viterbi, --no-viterbi, --fb  Use viterbi decoding only / forward-backward followed by viterbi\n"
          "      --uuid, --no-uuid      Output UUID / read file name\n"
          "      --weights=file         Weight bundle (default $FLAPPIE_B200_MODELS/<model>.ffbw)\n"
          "      --batch=nreads         Reads per device batch (default 1024)\n"
          "      --device=index         CUDA device (default 0)\n"
          "      --devices=list         CUDA devices to shard the reads over: \"0-7\", \"0,2,3\" or \"all\"\n"
          "      --stats                Print reads, samples and samples/s to stderr at the end\n", fh);
}

enum { OPT_SEG = 3, OPT_NOREV = 6, OPT_TEMP, OPT_NOVIT, OPT_FB, OPT_LIC, OPT_LIC2, OPT_H5C, OPT_H5K, OPT_UUID, OPT_NOUUID,
       OPT_WEIGHTS = 1000, OPT_BATCH, OPT_DEVICE, OPT_DEVICES, OPT_STATS, OPT_HELP };

static const struct option long_opts[] = {
    {"delta", required_argument, 0, 'd'}, {"format", required_argument, 0, 'f'}, {"limit", required_argument, 0, 'l'},
    {"model", required_argument, 0, 'm'}, {"output", required_argument, 0, 'o'}, {"prefix", required_argument, 0, 'p'},
    {"reverse", no_argument, 0, 'r'}, {"no-reverse", no_argument, 0, OPT_NOREV}, {"temperature", required_argument, 0, OPT_TEMP},
    {"trim", required_argument, 0, 't'}, {"trace", required_argument, 0, 'T'}, {"licence", no_argument, 0, OPT_LIC},
    {"license", no_argument, 0, OPT_LIC2}, {"segmentation", required_argument, 0, OPT_SEG}, {"viterbi", no_argument, 0, 'v'},
    {"no-viterbi", no_argument, 0, OPT_NOVIT}, {"fb", no_argument, 0, OPT_FB}, {"hdf5-compression", required_argument, 0, OPT_H5C},
    {"hdf5-chunk", required_argument, 0, OPT_H5K}, {"uuid", no_argument, 0, OPT_UUID}, {"no-uuid", no_argument, 0, OPT_NOUUID},
    {"weights", required_argument, 0, OPT_WEIGHTS}, {"batch", required_argument, 0, OPT_BATCH},
    {"device", required_argument, 0, OPT_DEVICE}, {"devices", required_argument, 0, OPT_DEVICES}, {"stats", no_argument, 0, OPT_STATS},
    {"help", no_argument, 0, OPT_HELP}, {0, 0, 0, 0}};

/* "0-7", "0,2,3", "1-2,5" or "all" */
static void parse_devices(const char *spec) {
    args.ndev = 0;
    if (0 == strcasecmp(spec, "all")) {
        const int n = ffb_device_count();
        for (int d = 0; d < n && d < MAX_DEVICES; d++) args.devices[args.ndev++] = d;
        if (args.ndev == 0) die("--devices all: no CUDA device%s", "");
        return;
    }
    char *copy = strdup(spec), *save = NULL;
    for (char *tok = strtok_r(copy, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1 || a < 0) die("--devices: cannot parse \"%s\"", spec);
        if (k == 1) b = a;
        for (int d = a; d <= b; d++) {
            if (args.ndev >= MAX_DEVICES) die("--devices: too many devices%s", "");
            args.devices[args.ndev++] = d;       /* naming a device twice gives it two independent pipelines */
        }
    }
    free(copy);
    if (args.ndev == 0) die("--devices: cannot parse \"%s\"", spec);
}

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
        case OPT_H5C: case OPT_H5K: break;   /* HDF5 tuning of the reference's --trace: nothing to tune in the flat file */
        case OPT_UUID: args.uuid = true; break;
        case OPT_NOUUID: args.uuid = false; break;
        case OPT_WEIGHTS: args.weights = optarg; break;
        case OPT_BATCH: args.batch = atoi(optarg) > 0 ? atoi(optarg) : 1; break;
        case OPT_DEVICE: args.ndev = 1; args.devices[0] = atoi(optarg); break;
        case OPT_DEVICES: parse_devices(optarg); break;
        case OPT_STATS: args.stats = true; break;
        case OPT_HELP: usage(stdout); exit(EXIT_SUCCESS);
        default: usage(stderr); exit(EXIT_FAILURE);
        }
    }
    if (optind >= argc) { usage(stderr); exit(EXIT_FAILURE); }
}

/* ---- the work list: every signal file named on the command line, in the reference's order ------------------------- */
struct filelist { char **path; size_t n, cap; };

static void filelist_add(struct filelist *fl, const char *p) {
    if (fl->n == fl->cap) {
        fl->cap = fl->cap ? 2 * fl->cap : 1024;
        fl->path = realloc(fl->path, sizeof(char *) * fl->cap);
        if (!fl->path) die("out of memory%s", "");
    }
    fl->path[fl->n++] = strdup(p);
}

/* files and directories, through the system glob as the reference does (src/flappie.c:338-362) */
static void expand_arguments(int argc, char **argv, struct filelist *fl) {
    for (int fn = optind; fn < argc; fn++) {
        if (args.limit > 0 && fl->n >= (size_t)args.limit) break;
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
            if (args.limit > 0 && fl->n >= (size_t)args.limit) break;
            filelist_add(fl, filename);
        }
        globfree(&globbuf);
    }
}

/* ---- one read of a window ------------------------------------------------------------------------------------------ */
struct read_slot {
    char *name, *uuid;      /* basename (printed as "filename", and as the id with --no-uuid); stem as the uuid */
    float *raw;             /* samples, malloc'ed by the reader, freed once packed */
    long n;                 /* > 0: samples; <= 0: unreadable / empty; -2: fast5 */
    int dev, pos;           /* device batch this read was dealt to, and its index there */
};

/* ---- one batch in flight on one context of one device; all result buffers are pinned and grow-only ------------------- */
struct dev_batch {
    ffb_ctx *ctx;
    int n;                          /* reads in this batch */
    int *member; int member_cap;    /* window-relative read index of each member, ascending */
    float *raw; size_t raw_cap;
    int64_t *raw_off, *blk_off, *start, *end; float *score; int32_t *nbases; size_t read_cap;
    char *bases, *quals; int32_t *path; float *qpath; float *rle; uint8_t *trace; size_t blk_cap;
    int64_t tot_blocks;
    ffb_batch b;
    ffb_raw_batch rb;
    bool busy;
};

/* what a device thread leaves behind for the printer once it has collected its batch of a window: the records as text
 * (read k of the batch = bytes [rec_off[k], rec_off[k+1])), the block counts, and -- with --trace -- a copy of the trace rows */
struct win_out { char *text; size_t text_len; size_t *rec_off; int64_t *nblock; uint8_t *trace; size_t *trace_off; int n; };
struct window { size_t first; int n; struct read_slot *rd; int cap; struct win_out out[MAX_DEVICES]; };

struct device {
    int id, rank;
    ffb_model *model;
    struct dev_batch bat[2];
    pthread_t th;
    double t_read, t_begin, t_finish, t_collect, t_format, t_wait;     /* --stats: seconds this thread spent where */
};

static struct {
    struct filelist files;
    struct window win[4];       /* window w lives in slot w % 4: while w is on the devices, w+1 is being read, w-1 collected
                                 * and formatted, w-2 printed */
    struct device dev[MAX_DEVICES];
    int ndev, nwin, nstate, nbase;
    size_t window_reads;
    pthread_barrier_t bar;
} G;

static void *pinned_grow(void *old, size_t bytes) {
    ffb_free_pinned(old);
    void *p = ffb_alloc_pinned(bytes);
    if (!p) die("out of pinned host memory%s", "");
    return p;
}

static void read_one(struct read_slot *s, const char *path) {
    char *tmp = strdup(path);
    s->name = strdup(basename(tmp));
    free(tmp);
    /* fast5 files carry a read uuid attribute; a bare signal file has none: its stem stands in */
    s->uuid = strdup(s->name);
    char *dot = strrchr(s->uuid, '.');
    if (dot) *dot = 0;
    s->raw = NULL;
    s->n = ffb_read_raw_file(path, &s->raw);
    if (s->n <= 0) { free(s->raw); s->raw = NULL; }
    s->dev = -1; s->pos = -1;
}

/* LPT deal of the window (ffb_shard.c), then every device's members in input order */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static void deal_window(struct window *w, int slot) {
    long *len = malloc(sizeof(long) * (size_t)(w->n > 0 ? w->n : 1));
    int *dev_of = malloc(sizeof(int) * (size_t)(w->n > 0 ? w->n : 1));
    if (!len || !dev_of) die("out of memory%s", "");
    for (int i = 0; i < w->n; i++) len[i] = w->rd[i].n;
    if (ffb_deal_lpt(len, w->n, G.ndev, args.batch, dev_of) < 0) die("internal: read sharding failed%s", "");
    for (int d = 0; d < G.ndev; d++) G.dev[d].bat[slot].n = 0;
    for (int i = 0; i < w->n; i++) {
        if (dev_of[i] < 0) continue;
        struct dev_batch *b = &G.dev[dev_of[i]].bat[slot];
        if (b->n == b->member_cap) {
            b->member_cap = b->member_cap ? 2 * b->member_cap : 1024;
            b->member = realloc(b->member, sizeof(int) * (size_t)b->member_cap);
            if (!b->member) die("out of memory%s", "");
        }
        b->member[b->n++] = i;
    }
    for (int d = 0; d < G.ndev; d++) {
        struct dev_batch *b = &G.dev[d].bat[slot];
        qsort(b->member, (size_t)b->n, sizeof(int), cmp_int);      /* input order within the batch */
        for (int k = 0; k < b->n; k++) { w->rd[b->member[k]].dev = d; w->rd[b->member[k]].pos = k; }
    }
    free(len); free(dev_of);
}

/* grow-only pinned buffers of one batch slot: reads, raw samples, blocks (+ one entry per read) */
static void reserve_batch(struct dev_batch *f, size_t n, size_t tot_raw, size_t need_blk) {
    if (n > f->read_cap) {
        f->read_cap = n + n / 4 + 16;
        f->raw_off = pinned_grow(f->raw_off, (f->read_cap + 1) * sizeof(int64_t));
        f->blk_off = pinned_grow(f->blk_off, (f->read_cap + 1) * sizeof(int64_t));
        f->start = pinned_grow(f->start, f->read_cap * sizeof(int64_t));
        f->end = pinned_grow(f->end, f->read_cap * sizeof(int64_t));
        f->score = pinned_grow(f->score, f->read_cap * sizeof(float));
        f->nbases = pinned_grow(f->nbases, f->read_cap * sizeof(int32_t));
    }
    if (tot_raw > f->raw_cap) {
        f->raw_cap = tot_raw + tot_raw / 4 + 4096;
        f->raw = pinned_grow(f->raw, f->raw_cap * sizeof(float));
    }
    if (need_blk > f->blk_cap) {
        f->blk_cap = need_blk + need_blk / 4 + 4096;
#ifdef FFB_RUNNIE
        f->path = pinned_grow(f->path, f->blk_cap * sizeof(int32_t));
        f->qpath = pinned_grow(f->qpath, f->blk_cap * sizeof(float));
        f->rle = pinned_grow(f->rle, f->blk_cap * 8 * sizeof(float));
#else
        f->bases = pinned_grow(f->bases, f->blk_cap);
        f->quals = pinned_grow(f->quals, f->blk_cap);
        if (args.trace) f->trace = pinned_grow(f->trace, f->blk_cap * (size_t)G.nstate);
#endif
    }
}

/* pack the batch into pinned memory and enqueue its raw upload + trimming (ffb_submit_raw_begin: returns at once) */
static void submit_begin(struct device *dv, struct dev_batch *f, struct window *w) {
    f->busy = false;
    if (f->n == 0) return;
    const int n = f->n;
    size_t tot_raw = 0;
    int64_t tot_blocks = 0;
    for (int k = 0; k < n; k++) {
        const long len = w->rd[f->member[k]].n;
        tot_raw += (size_t)len;
        const long t = ffb_model_nblock(dv->model, len);
        tot_blocks += t > 0 ? t : 0;                  /* upper bound: the kept range is shorter */
    }
    reserve_batch(f, (size_t)n, tot_raw, (size_t)tot_blocks + (size_t)n + 1);
    f->raw_off[0] = 0;
    for (int k = 0; k < n; k++) {
        struct read_slot *s = &w->rd[f->member[k]];
        memcpy(f->raw + f->raw_off[k], s->raw, sizeof(float) * (size_t)s->n);
        f->raw_off[k + 1] = f->raw_off[k] + s->n;
        free(s->raw);
        s->raw = NULL;
    }
    f->tot_blocks = tot_blocks;
    f->rb = (ffb_raw_batch){.raw = f->raw, .raw_off = f->raw_off, .n_reads = n, .trim_start = args.trim_start, .trim_end = args.trim_end,
                            .varseg_chunk = args.varseg_chunk, .varseg_thresh = args.varseg_thresh, .delta = args.delta,
                            .start = f->start, .end = f->end};
    memset(&f->b, 0, sizeof f->b);
    f->b.n_reads = n; f->b.temperature = args.temperature;
    f->b.flags = (args.viterbi_only ? FFB_FLAG_VITERBI_ONLY : 0) | (args.reverse ? FFB_FLAG_REVERSE : 0) | (args.trace ? FFB_FLAG_WANT_TRACE : 0);
    f->b.blk_off = f->blk_off; f->b.score = f->score;
#ifdef FFB_RUNNIE
    f->b.path = f->path; f->b.qpath = f->qpath; f->b.rle_params = f->rle;
#else
    /* bases and quality characters come off the device (emit.cu): path / qpath stay there */
    f->b.bases = f->bases; f->b.quals = f->quals; f->b.nbases = f->nbases; f->b.trace = f->trace;
#endif
    if (ffb_submit_raw_begin(f->ctx, &f->rb, &f->b) != FFB_OK) die("device basecall failed: %s", ffb_last_error());
    f->busy = true;
}

/* ... and, once the trim bounds are back, the plan and every other kernel (ffb_submit_raw_finish) */
static void submit_finish(struct dev_batch *f) {
    if (!f->busy) return;
    if (ffb_submit_raw_finish(f->ctx) != FFB_OK) die("device basecall failed: %s", ffb_last_error());
}

static void collect_batch(struct dev_batch *f) {
    if (!f->busy) return;
    if (ffb_collect(f->ctx, &f->b) != FFB_OK) die("device basecall failed: %s", ffb_last_error());
    f->busy = false;
}

/* the reference's per-read printing (src/flappie.c:364-385 -> fprintf_format), into memory: one device thread per batch */
static void format_batch(const struct dev_batch *f, struct window *w, int d) {
    struct win_out *o = &w->out[d];
    memset(o, 0, sizeof *o);
    o->n = f->n;
    if (f->n == 0) return;
    o->rec_off = malloc(sizeof(size_t) * ((size_t)f->n + 1));
    o->nblock = malloc(sizeof(int64_t) * (size_t)f->n);
    if (!o->rec_off || !o->nblock) die("out of memory%s", "");
    FILE *ms = open_memstream(&o->text, &o->text_len);
    if (!ms) die("out of memory%s", "");
    o->rec_off[0] = 0;
    size_t trace_bytes = 0;
    for (int k = 0; k < f->n; k++) {
        const struct read_slot *s = &w->rd[f->member[k]];
        const int64_t nblock = f->blk_off[k + 1] - f->blk_off[k];
        o->nblock[k] = nblock;
        if (nblock > 0) {
            const int64_t o0 = f->blk_off[k] + k;
            trace_bytes += (size_t)(nblock + 1) * (size_t)G.nstate;
#ifdef FFB_RUNNIE
            char *bases = calloc((size_t)nblock + 2, 1);
            float *shape = calloc((size_t)nblock + 1, sizeof(float)), *scale = calloc((size_t)nblock + 1, sizeof(float));
            int32_t *dwell = calloc((size_t)nblock + 1, sizeof(int32_t));
            if (!bases || !shape || !scale || !dwell) die("out of memory%s", "");
            const int64_t nrun = ffb_emit_runs(f->path + o0, f->rle + f->blk_off[k] * 8, nblock, G.nbase, bases, shape, scale, dwell);
            fprintf(ms, "# %s\n", s->uuid);                                  /* src/runnie.c:277 */
            for (int64_t r = 0; r < nrun; r++) fprintf(ms, "%c\t%f\t%f\t%d\n", bases[r], shape[r], scale[r], dwell[r]);
            free(bases); free(shape); free(scale); free(dwell);
#else
            ffb_read_result res = {.score = f->score[k], .n = (size_t)s->n, .start = (size_t)f->start[k], .end = (size_t)f->end[k],
                                   .basecall = f->bases + o0, .quality = f->quals + o0, .basecall_length = (size_t)f->nbases[k],
                                   .nblock = (size_t)nblock};
            ffb_fprintf_read(args.outformat, ms, s->uuid, s->name, args.uuid, args.prefix, &res);
#endif
        }
        fflush(ms);
        o->rec_off[k + 1] = o->text_len;
    }
    fclose(ms);
#ifndef FFB_RUNNIE
    if (args.trace && trace_bytes > 0) {          /* the pinned trace buffer is reused two windows later: keep the rows */
        o->trace = malloc(trace_bytes);
        o->trace_off = malloc(sizeof(size_t) * ((size_t)f->n + 1));
        if (!o->trace || !o->trace_off) die("out of memory%s", "");
        size_t at = 0;
        for (int k = 0; k < f->n; k++) {
            o->trace_off[k] = at;
            if (o->nblock[k] > 0) {
                const size_t nb = (size_t)(o->nblock[k] + 1) * (size_t)G.nstate;
                memcpy(o->trace + at, f->trace + (size_t)(f->blk_off[k] + k) * (size_t)G.nstate, nb);
                at += nb;
            }
        }
        o->trace_off[f->n] = at;
    }
#endif
}

/* one window, in input order: the records the device threads formatted, the reference's messages for the reads without one */
static FILE *trace_fp = NULL;
static void print_window(struct window *w, int64_t *reads_called, int64_t *samples) {
    for (int i = 0; i < w->n; i++) {
        struct read_slot *s = &w->rd[i];
        if (s->n == -2) fprintf(stderr, PROGRAM ": %s: fast5 input needs libhdf5, which this build does not have\n", s->name);
        const struct win_out *o = s->dev >= 0 ? &w->out[s->dev] : NULL;
        const int k = s->pos;
        const int64_t nblock = o ? o->nblock[k] : 0;
        if (nblock <= 0) {
            if (s->n != -2) fprintf(stderr, PROGRAM ": No basecall returned for %s\n", s->name);   /* src/flappie.c:370-373 */
            free(s->name); free(s->uuid);
            continue;
        }
        *reads_called += 1;
        *samples += s->n;
        fwrite(o->text + o->rec_off[k], 1, o->rec_off[k + 1] - o->rec_off[k], args.output);
        if (trace_fp && o->trace) ffb_write_trace(trace_fp, args.uuid ? s->uuid : s->name, o->trace + o->trace_off[k], (size_t)nblock, (size_t)G.nstate);
        free(s->name); free(s->uuid);
    }
    fflush(args.output);
    for (int d = 0; d < G.ndev; d++) {
        struct win_out *o = &w->out[d];
        free(o->text); free(o->rec_off); free(o->nblock); free(o->trace); free(o->trace_off);
        memset(o, 0, sizeof *o);
    }
}

/* ---- device thread.  Iteration w: enqueue the raw upload + trimming of window w, read the files of window w+1 while the
 * device trims (and still computes window w-1), plan and enqueue the rest of w, then collect and format window w-1 ------ */
static void read_window(const struct device *dv, struct window *win) {
    for (int i = dv->rank; i < win->n; i += G.ndev) read_one(&win->rd[i], G.files.path[win->first + (size_t)i]);
}

static double now_s(void);
#define TIMED(acc, stmt) do { const double t_ = now_s(); stmt; (acc) += now_s() - t_; } while (0)

static void *device_main(void *arg) {
    struct device *dv = arg;
    if (G.nwin > 0) TIMED(dv->t_read, read_window(dv, &G.win[0]));
    for (int w = 0; w <= G.nwin; w++) {
        TIMED(dv->t_wait, pthread_barrier_wait(&G.bar));       /* B1: window w is in memory */
        TIMED(dv->t_wait, pthread_barrier_wait(&G.bar));       /* B2: the main thread has dealt it (and set the geometry of window w+1) */
        if (w < G.nwin) TIMED(dv->t_begin, submit_begin(dv, &dv->bat[w & 1], &G.win[w % 4]));
        if (w + 1 < G.nwin) TIMED(dv->t_read, read_window(dv, &G.win[(w + 1) % 4]));
        if (w < G.nwin) TIMED(dv->t_finish, submit_finish(&dv->bat[w & 1]));
        if (w > 0) {
            TIMED(dv->t_collect, collect_batch(&dv->bat[(w - 1) & 1]));
            TIMED(dv->t_format, format_batch(&dv->bat[(w - 1) & 1], &G.win[(w - 1) % 4], dv->rank));
        }
        TIMED(dv->t_wait, pthread_barrier_wait(&G.bar));       /* B3: window w-1 is formatted; the main thread prints it during iteration w+1 */
    }
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
    parse_args(argc, argv);
    if (!args.output) args.output = stdout;
    setvbuf(args.output, NULL, _IOFBF, 1 << 22);
#ifdef FFB_RUNNIE
    if (args.trace) die("--trace %s: runnie writes no trace", args.trace);
#endif
    const int have = ffb_device_count();
    for (int d = 0; d < args.ndev; d++)
        if (args.devices[d] < 0 || args.devices[d] >= have) die("no such CUDA device (flappie_b200 has no CPU fallback)%s", "");

    char wpath[4096];
    if (!args.weights) {
        const char *dir = getenv("FLAPPIE_B200_MODELS");
        if (!dir) die("no weights: give --weights <bundle> or set FLAPPIE_B200_MODELS (model %s)", args.model_name);
        snprintf(wpath, sizeof wpath, "%s/%s.ffbw", dir, args.model_name);
        args.weights = wpath;
    }
    ffb_bundle bundle;
    if (ffb_bundle_load(args.weights, &bundle) != 0) die("cannot read weight bundle \"%s\"", args.weights);
    G.ndev = args.ndev;
    for (int d = 0; d < G.ndev; d++) {
        struct device *dv = &G.dev[d];
        dv->id = args.devices[d]; dv->rank = d;
        dv->model = ffb_bundle_to_model(&bundle, dv->id);
        if (!dv->model) die("weight bundle rejected: %s", ffb_last_error());
        for (int k = 0; k < 2; k++) {
            dv->bat[k].ctx = ffb_create(dv->model, NULL);
            if (!dv->bat[k].ctx) die("ffb_create: %s", ffb_last_error());
        }
    }
    ffb_bundle_free(&bundle);
    G.nbase = (int)nbase_from_flipflop_nparam((size_t)ffb_model_nparam(G.dev[0].model));
    G.nstate = 2 * G.nbase;
    /* start-up, like the weight upload: page-locked buffers for batches of --batch reads of up to 8192 samples each (they
     * still grow when the reads turn out longer); cudaHostAlloc is slow and serialises across threads, so not in the loop */
    for (int d = 0; d < G.ndev; d++)
        for (int k = 0; k < 2; k++) {
            const long nb = ffb_model_nblock(G.dev[d].model, 8192);
            reserve_batch(&G.dev[d].bat[k], (size_t)args.batch, (size_t)args.batch * 8192, (size_t)args.batch * (size_t)((nb > 0 ? nb : 4096) + 1) + 1);
            /* ... and the context's device workspaces for batches of 4096-sample reads (grow-only as well) */
            if (ffb_reserve(G.dev[d].bat[k].ctx, args.batch, 4096, (args.viterbi_only ? FFB_FLAG_VITERBI_ONLY : 0) | (args.trace ? FFB_FLAG_WANT_TRACE : 0)) != FFB_OK)
                die("ffb_reserve: %s", ffb_last_error());
        }
    if (args.trace) {
        trace_fp = fopen(args.trace, "wb");
        if (!trace_fp) die("Failed to open \"%s\" for the trace.", args.trace);
    }

    const double t0 = now_s();
    expand_arguments(argc, argv, &G.files);
    G.window_reads = (size_t)G.ndev * (size_t)args.batch;
    G.nwin = (int)((G.files.n + G.window_reads - 1) / G.window_reads);
    for (int k = 0; k < 4; k++) {
        G.win[k].cap = (int)G.window_reads;
        G.win[k].rd = calloc(G.window_reads, sizeof(struct read_slot));
        if (!G.win[k].rd) die("out of memory%s", "");
    }
    pthread_barrier_init(&G.bar, NULL, (unsigned)G.ndev + 1);
    /* window geometry is a pure function of w; slot w % 4 is set by the main thread before any device thread reads it:
     * window 0 here, window w+1 between B1 and B2 of iteration w (its slot held window w-3, printed two iterations ago) */
    if (G.nwin > 0) {
        G.win[0].first = 0;
        G.win[0].n = (int)(G.files.n < G.window_reads ? G.files.n : G.window_reads);
    }
    for (int d = 0; d < G.ndev; d++)
        if (pthread_create(&G.dev[d].th, NULL, device_main, &G.dev[d]) != 0) die("pthread_create failed%s", "");

    int64_t reads_called = 0, samples = 0;
    double m_deal = 0, m_print = 0, m_wait = 0;
    const double t_list = now_s() - t0;
    for (int w = 0; w <= G.nwin; w++) {
        TIMED(m_wait, pthread_barrier_wait(&G.bar));       /* B1 */
        if (w < G.nwin) TIMED(m_deal, deal_window(&G.win[w % 4], w & 1));
        if (w + 1 < G.nwin) {
            struct window *nx = &G.win[(w + 1) % 4];
            nx->first = (size_t)(w + 1) * G.window_reads;
            const size_t left = G.files.n - nx->first;
            nx->n = (int)(left < G.window_reads ? left : G.window_reads);
        }
        TIMED(m_wait, pthread_barrier_wait(&G.bar));       /* B2 */
        /* while the device threads work on iteration w: print window w-2 (formatted before B3 of iteration w-1) */
        if (w >= 2) TIMED(m_print, print_window(&G.win[(w - 2) % 4], &reads_called, &samples));
        TIMED(m_wait, pthread_barrier_wait(&G.bar));       /* B3 */
    }
    if (G.nwin >= 1) TIMED(m_print, print_window(&G.win[(G.nwin - 1) % 4], &reads_called, &samples));     /* formatted in the last iteration */
    for (int d = 0; d < G.ndev; d++) pthread_join(G.dev[d].th, NULL);
    const double t1 = now_s();
    if (args.stats)
        fprintf(stderr, PROGRAM ": stats { \"devices\": %d, \"files\": %zu, \"reads_called\": %lld, \"samples\": %lld, \"seconds\": %.3f, "
                "\"samples_per_s\": %.0f }\n", G.ndev, G.files.n, (long long)reads_called, (long long)samples, t1 - t0,
                (double)samples / (t1 - t0 > 0 ? t1 - t0 : 1));
    if (args.stats) {
        fprintf(stderr, PROGRAM ": where the host time went (s): file list %.3f; main thread: deal %.3f print %.3f waiting %.3f\n", t_list, m_deal, m_print, m_wait);
        for (int d = 0; d < G.ndev; d++)
            fprintf(stderr, PROGRAM ":   device thread %d (GPU %d): read files %.3f  pack+begin %.3f  plan+enqueue %.3f  wait for GPU %.3f  format %.3f  barriers %.3f\n",
                    d, G.dev[d].id, G.dev[d].t_read, G.dev[d].t_begin, G.dev[d].t_finish, G.dev[d].t_collect, G.dev[d].t_format, G.dev[d].t_wait);
    }

    for (int d = 0; d < G.ndev; d++) {
        for (int k = 0; k < 2; k++) {
            struct dev_batch *f = &G.dev[d].bat[k];
            ffb_destroy(f->ctx);
            ffb_free_pinned(f->raw); ffb_free_pinned(f->raw_off); ffb_free_pinned(f->blk_off); ffb_free_pinned(f->start);
            ffb_free_pinned(f->end); ffb_free_pinned(f->score); ffb_free_pinned(f->nbases); ffb_free_pinned(f->bases);
            ffb_free_pinned(f->quals); ffb_free_pinned(f->path); ffb_free_pinned(f->qpath); ffb_free_pinned(f->rle);
            ffb_free_pinned(f->trace);
            free(f->member);
        }
        ffb_model_destroy(G.dev[d].model);
    }
    for (size_t i = 0; i < G.files.n; i++) free(G.files.path[i]);
    free(G.files.path); free(G.win[0].rd); free(G.win[1].rd); free(G.win[2].rd); free(G.win[3].rd);
    if (trace_fp) fclose(trace_fp);
    if (stdout != args.output) fclose(args.output);
    return EXIT_SUCCESS;
}
