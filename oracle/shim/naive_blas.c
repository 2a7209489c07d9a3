/* Fallback BLAS for oracle/_ref when no OpenBLAS binary can be found: only the
 * two entry points and the argument combinations the reference uses
 * (ColMajor; sgemm Trans/NoTrans; sgemv Trans; unit increments).
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library. */
#include <stddef.h>
#include <stdlib.h>
#include "cblas.h"

void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta,
                 const enum CBLAS_TRANSPOSE tb, const int m, const int n, const int k,
                 const float alpha, const float *a, const int lda, const float *b,
                 const int ldb, const float beta, float *c, const int ldc) {
    if (order != CblasColMajor || ta != CblasTrans || tb != CblasNoTrans) abort();
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) {
            const float *ai = a + (size_t)i * lda;
            const float *bj = b + (size_t)j * ldb;
            float acc = 0.0f;
            for (int l = 0; l < k; l++) acc += ai[l] * bj[l];
            float *cij = c + (size_t)j * ldc + i;
            *cij = alpha * acc + beta * (*cij);
        }
    }
}

void cblas_sgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int m,
                 const int n, const float alpha, const float *a, const int lda,
                 const float *x, const int incx, const float beta, float *y, const int incy) {
    if (order != CblasColMajor || ta != CblasTrans) abort();
    for (int j = 0; j < n; j++) {
        const float *aj = a + (size_t)j * lda;
        float acc = 0.0f;
        for (int l = 0; l < m; l++) acc += aj[l] * x[(size_t)l * incx];
        y[(size_t)j * incy] = alpha * acc + beta * y[(size_t)j * incy];
    }
}
void openblas_set_num_threads(int n) { (void)n; }
