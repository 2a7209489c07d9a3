/* ffb_shard.c -- dealing the reads of a window to the GPUs of the box (reads shard embarrassingly, the
 * partition IS the multi-GPU design; flappie_b200/shard.py:shard_reads is the same rule for Python callers).
 * Longest read first, each to the device with the fewest samples so far that still has room: equal work per device
 * whatever the length distribution (BASELINE configs[3]: 1 k - 50 k samples).  Deterministic: ties go to the lower
 * read index / lower device. */
#define _GNU_SOURCE
#include <stdlib.h>

#include "ffb_host.h"

static int cmp_len_desc(const void *a, const void *b, void *ctx) {
    const long *len = ctx;
    const long la = len[*(const int *)a], lb = len[*(const int *)b];
    if (la != lb) return la > lb ? -1 : 1;
    return *(const int *)a - *(const int *)b;
}

int ffb_deal_lpt(const long *len, int n, int ndev, int cap, int *dev_of) {
    if (!len || !dev_of || n < 0 || ndev < 1 || cap < 1) return -1;
    int *idx = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    long long *load = calloc((size_t)ndev, sizeof(long long));
    int *count = calloc((size_t)ndev, sizeof(int));
    if (!idx || !load || !count) { free(idx); free(load); free(count); return -1; }
    int m = 0;
    for (int i = 0; i < n; i++) {
        dev_of[i] = -1;
        if (len[i] > 0) idx[m++] = i;
    }
    qsort_r(idx, (size_t)m, sizeof(int), cmp_len_desc, (void *)len);
    int dealt = 0;
    for (int k = 0; k < m; k++) {
        int best = -1;
        for (int d = 0; d < ndev; d++) {
            if (count[d] >= cap) continue;
            if (best < 0 || load[d] < load[best]) best = d;
        }
        if (best < 0) break;                 /* every device is full: the caller's window was too large */
        dev_of[idx[k]] = best;
        load[best] += len[idx[k]];
        count[best] += 1;
        dealt++;
    }
    free(idx); free(load); free(count);
    return dealt;
}
