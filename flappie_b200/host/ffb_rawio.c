/* ffb_rawio.c -- raw signal readers for the command line (stand-ins for read_raw, src/fast5_interface.c:231-300,
 * which needs libhdf5). */
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ffb_host.h"

static bool has_suffix(const char *s, const char *suf) {
    const size_t n = strlen(s), m = strlen(suf);
    return n >= m && 0 == strcmp(s + n - m, suf);
}

bool ffb_is_signal_file(const char *path) {
    return path && (has_suffix(path, ".f32") || has_suffix(path, ".crp") || has_suffix(path, ".fast5"));
}

/* one open / fstat / read / close per file: the command line reads thousands of single-read files per second per thread */
static long read_f32(const char *path, float **out) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat sb;
    long n = -1;
    if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size % (off_t)sizeof(float) == 0) {
        n = (long)(sb.st_size / (off_t)sizeof(float));
        float *buf = malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
        size_t got = 0;
        const size_t want = (size_t)n * sizeof(float);
        while (buf && got < want) {
            const ssize_t r = read(fd, (char *)buf + got, want - got);
            if (r <= 0) break;
            got += (size_t)r;
        }
        if (!buf || got != want) { free(buf); n = -1; }
        else *out = buf;
    }
    close(fd);
    return n;
}

/* text matrix "nr<TAB>nc" followed by one line per COLUMN with nr hex-float entries (what
 * write_flappie_matrix_to_handle emits, src/test/flappie_util.c:30-55); the signal fixtures are 1 x T, so the
 * signal is row 0 */
static long read_crp(const char *path, float **out) {
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    long nr = 0, nc = 0;
    if (fscanf(fp, "%ld %ld", &nr, &nc) != 2 || nr < 1 || nc < 0) { fclose(fp); return -1; }
    float *buf = malloc((size_t)(nc > 0 ? nc : 1) * sizeof(float));
    if (!buf) { fclose(fp); return -1; }
    for (long c = 0; c < nc; c++)
        for (long r = 0; r < nr; r++) {
            float v;
            if (fscanf(fp, "%f", &v) != 1) { free(buf); fclose(fp); return -1; }
            if (r == 0) buf[c] = v;
        }
    fclose(fp);
    *out = buf;
    return nc;
}

long ffb_read_raw_file(const char *path, float **out) {
    if (!path || !out) return -1;
    *out = NULL;
    if (has_suffix(path, ".fast5")) return -2;
    if (has_suffix(path, ".crp")) return read_crp(path, out);
    return read_f32(path, out);
}
