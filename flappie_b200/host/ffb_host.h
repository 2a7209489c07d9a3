/* ffb_host.h -- C99 host side of the flappie_b200 command line: output writers, weight-bundle loader and
 * raw-signal readers.  Everything numerical happens behind include/flappie_b200.h (CUDA); this is the part of
 * the reference that "stays C" (src/flappie.c main loop, src/flappie_output.c, src/fast5_interface.c:read_raw). */
#ifndef FFB_HOST_H
#define FFB_HOST_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/flappie_b200.h"

/* ---- output formats (reference src/flappie_output.h:13-19, src/flappie_output.c) ---- */
enum ffb_outformat { FFB_OUT_FASTA = 0, FFB_OUT_FASTQ, FFB_OUT_SAM, FFB_OUT_INVALID };
enum ffb_outformat ffb_get_outformat(const char *name);              /* "fasta" | "fastq" | "sam" */
const char *ffb_outformat_string(enum ffb_outformat fmt);            /* NULL for an invalid value */

/* What calculate_post hands to the writers (reference struct _raw_basecall_info, src/flappie_structures.h:24-36) */
typedef struct {
    float score;
    size_t n, start, end;        /* raw_table: untrimmed length and kept range */
    const char *basecall;
    const char *quality;         /* may be NULL (fastq then refuses, as the reference does) */
    size_t basecall_length;
    size_t nblock;
} ffb_read_result;

/* One record in `fmt`, byte for byte what the reference's fprintf_format prints (including SAM's second,
 * header-less line -- src/flappie_output.c:123-133). */
void ffb_fprintf_read(enum ffb_outformat fmt, FILE *fp, const char *uuid, const char *readname, bool uuid_primary,
                      const char *prefix, const ffb_read_result *res);

/* ---- state trace (--trace) -----------------------------------------------------------------
 * The reference writes the trace of every read into an HDF5 file (src/fast5_interface.c:126-143, one u8 dataset
 * [nblock + 1][nstate] per read).  Without libhdf5 the same bytes go into a flat file of records
 *   char magic[4] = "FFBT"; uint32 name_len; char name[name_len]; uint64 nrow (= nblock + 1); uint32 nstate;
 *   uint8 trace[nrow * nstate]          (row-major: row = block boundary, column = flip-flop state)
 * in input order. */
void ffb_write_trace(FILE *fp, const char *name, const uint8_t *trace, size_t nblock, size_t nstate);

/* ---- weight bundles -----------------------------------------------------------------
 * The reference compiles its weights in from generated headers (src/models/ *.mdl, git-LFS); this driver loads the
 * same `_Mat` images from a binary bundle instead:
 *   char magic[8] = "FFBW1\0\0\0"; int32 kind; int32 nconv; int32 stride[3]; int32 nmat;
 *   nmat x { uint64 nr; uint64 nc; float data[nc * 4*ceil(nr/4)] }      (column-major, zero-padded columns)
 * in the field order of guppy_model / guppy_stride5_model (src/networks.c:150-215). */
typedef struct {
    int kind, nconv, nmat;
    int stride[3];
    _Mat *mats;                  /* nmat matrices, data owned by the bundle */
} ffb_bundle;
int ffb_bundle_load(const char *path, ffb_bundle *out);   /* 0 on success */
void ffb_bundle_free(ffb_bundle *b);
ffb_model *ffb_bundle_to_model(const ffb_bundle *b, int device);

/* ---- read sharding over the GPUs of the box ----------------------------------------------
 * len[i] = samples of read i (<= 0: unreadable, dealt to nobody); dev_of[i] = device rank in [0, ndev) or -1.  Longest
 * first to the least-loaded device with fewer than `cap` reads.  Returns the number of reads dealt, -1 on bad arguments. */
int ffb_deal_lpt(const long *len, int n, int ndev, int cap, int *dev_of);

/* ---- raw signal input -----------------------------------------------------------------
 * <name>.f32 : little-endian float32 samples in pA, as read_raw(..., scale=true) returns them
 *              (src/fast5_interface.c:231-300)
 * <name>.crp : the reference's text matrix format (src/flappie_util.c:30-132), first column = samples
 * <name>.fast5 needs libhdf5, which this build does not have: returns -2.
 * Returns the number of samples (>= 0) with *out malloc'ed, or a negative error. */
long ffb_read_raw_file(const char *path, float **out);
bool ffb_is_signal_file(const char *path);

#endif
