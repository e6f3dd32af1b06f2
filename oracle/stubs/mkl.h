/*
 * Minimal stand-in for Intel MKL, used ONLY to compile the untouched reference headers
 * into oracle/_ref/. The reference calls MKL for the fp32 sgemv / somatcopy and an init
 * banner (CloverMatrix32.h:106,126,192; CloverBase.h:352-362) - none of it on the
 * 4/8-bit hot path. TEST INFRASTRUCTURE - not product code.
 */
#ifndef CLOVER_B200_ORACLE_MKL_STUB_H
#define CLOVER_B200_ORACLE_MKL_STUB_H
#include <stddef.h>
#include <string.h>
enum CBLAS_LAYOUT { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112 };
static inline void mkl_get_version_string(char *buf, int len) { strncpy(buf, "mkl-stub (clover_b200 oracle)", (size_t)len); }
static inline void mkl_set_num_threads(int) {}
static inline void cblas_sgemv(int, int, int m, int n, float alpha, const float *A, int lda,
                               const float *x, int incx, float beta, float *y, int incy) {
    for (int i = 0; i < m; ++i) {
        double acc = 0;
        for (int j = 0; j < n; ++j) acc += (double)A[(size_t)i * lda + j] * x[(size_t)j * incx];
        y[(size_t)i * incy] = alpha * (float)acc + (beta == 0.0f ? 0.0f : beta * y[(size_t)i * incy]);
    }
}
static inline void mkl_somatcopy(char, char, size_t rows, size_t cols, float alpha, const float *A, size_t lda,
                                 float *B, size_t ldb) {
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) B[j * ldb + i] = alpha * A[i * lda + j];
}
#endif
