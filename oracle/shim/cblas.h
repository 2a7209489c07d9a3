/* Minimal CBLAS declarations for building the reference's layers.c /
 * flappie_matrix.c (they `#include <cblas.h>`, reference src/layers.c:12,
 * src/flappie_matrix.c:12) in an image that ships OpenBLAS binaries (inside
 * pip wheels) but no BLAS headers.  TEST INFRASTRUCTURE ONLY. */
#ifndef FFB_SHIM_CBLAS_H
#define FFB_SHIM_CBLAS_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta,
                 const enum CBLAS_TRANSPOSE tb, const int m, const int n, const int k,
                 const float alpha, const float *a, const int lda, const float *b,
                 const int ldb, const float beta, float *c, const int ldc);
void cblas_sgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int m,
                 const int n, const float alpha, const float *a, const int lda,
                 const float *x, const int incx, const float beta, float *y, const int incy);
#endif
