/* oracle/linkproof/fast5_stub.c -- TEST INFRASTRUCTURE.
 * Stands in for the reference's src/fast5_interface.c (libhdf5, absent from this image) in the link proof: the three
 * functions src/fast5_interface.h declares.  read_raw returns what the reference's returns for a single-read fast5 with
 * scale_to_pA = true (src/fast5_interface.c:231-300): samples in pA, start = 0, end = n, a malloc'ed uuid -- read here
 * from a little-endian float32 file <name>.f32 whose stem is the uuid. */
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fast5_interface.h"

raw_table read_raw(const char *filename, bool scale_to_pA) {
    raw_table rt = {NULL, 0, 0, 0, NULL};
    (void)scale_to_pA;
    FILE *fp = fopen(filename, "rb");
    if (!fp) { fprintf(stderr, "Failed to open %s for reading.\n", filename); return rt; }
    fseek(fp, 0, SEEK_END);
    const long bytes = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    const size_t n = bytes > 0 ? (size_t)bytes / sizeof(float) : 0;
    float *raw = calloc(n ? n : 1, sizeof(float));
    if (!raw || fread(raw, sizeof(float), n, fp) != n) { free(raw); fclose(fp); return rt; }
    fclose(fp);
    char *tmp = strdup(filename);
    char *stem = strdup(basename(tmp));
    free(tmp);
    char *dot = strrchr(stem, '.');
    if (dot) *dot = 0;
    rt.uuid = stem; rt.n = n; rt.start = 0; rt.end = n; rt.raw = raw;
    return rt;
}

hid_t open_or_create_hdf5(const char *filename) { (void)filename; return -1; }

void write_summary(hid_t hdf5file, const char *readname, const struct _raw_basecall_info res, hsize_t chunk_size,
                   int compression_level) {
    (void)hdf5file; (void)readname; (void)res; (void)chunk_size; (void)compression_level;
}
