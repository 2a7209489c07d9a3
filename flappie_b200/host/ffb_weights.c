/* ffb_weights.c -- load a weight bundle (`_Mat` images in the reference's struct order) and hand it to the device. */
#include <stdlib.h>
#include <string.h>

#include "ffb_host.h"

static const char MAGIC[8] = {'F', 'F', 'B', 'W', '1', 0, 0, 0};

void ffb_bundle_free(ffb_bundle *b) {
    if (!b) return;
    if (b->mats) {
        for (int i = 0; i < b->nmat; i++) free(b->mats[i].data.v);
        free(b->mats);
    }
    memset(b, 0, sizeof *b);
}

int ffb_bundle_load(const char *path, ffb_bundle *out) {
    if (!path || !out) return -1;
    memset(out, 0, sizeof *out);
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    char magic[8];
    int32_t head[6];
    int rc = -1;
    if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, MAGIC, 8) != 0) goto done;
    if (fread(head, sizeof(int32_t), 6, fp) != 6) goto done;
    out->kind = head[0]; out->nconv = head[1];
    out->stride[0] = head[2]; out->stride[1] = head[3]; out->stride[2] = head[4];
    out->nmat = head[5];
    if (out->nmat < 1 || out->nmat > 64 || out->nconv < 1 || out->nconv > 3) goto done;
    out->mats = calloc((size_t)out->nmat, sizeof(_Mat));
    if (!out->mats) goto done;
    for (int i = 0; i < out->nmat; i++) {
        uint64_t dim[2];
        if (fread(dim, sizeof(uint64_t), 2, fp) != 2 || dim[0] == 0 || dim[1] == 0 || dim[0] > (1u << 24) || dim[1] > (1u << 24)) goto done;
        _Mat *m = &out->mats[i];
        m->nr = (size_t)dim[0]; m->nc = (size_t)dim[1];
        m->nrq = (m->nr + 3) / 4; m->stride = 4 * m->nrq;
        const size_t nfloat = m->stride * m->nc;
        void *p = NULL;
        if (posix_memalign(&p, 16, nfloat * sizeof(float)) != 0) goto done;   /* as make_flappie_matrix does */
        m->data.v = p;
        if (fread(p, sizeof(float), nfloat, fp) != nfloat) goto done;
    }
    rc = 0;
done:
    fclose(fp);
    if (rc != 0) ffb_bundle_free(out);
    return rc;
}

ffb_model *ffb_bundle_to_model(const ffb_bundle *b, int device) {
    if (!b || !b->mats) return NULL;
    const _Mat *ptr[64];
    for (int i = 0; i < b->nmat; i++) ptr[i] = &b->mats[i];
    return ffb_model_create(device, b->kind, ptr, b->nmat, b->stride, b->nconv);
}
