/*
 * oracle/clover_oracle.h - CPU restatement of the reference algorithms on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the CHECKER. The product path (clover_b200/, include/) never links it.
 *
 * Parity status: PINNED. Every function here is checked bit-for-bit against the unmodified
 * reference compiled from /root/reference (oracle/_ref, tests/test_oracle_vs_reference.py)
 * and against the committed golden vectors generated from it (tests/golden/).
 *
 * Conventions shared by all entry points:
 *   - `n` is the logical vector length; buffers are sized for n_pad = orc_size_pad(n)
 *     (reference: include/CloverVector.h:86-89). fp32 inputs must have their pad zeroed.
 *   - PRNG state = uint64_t[8] = part1[4] | part2[4] (include/CloverRandom.h:39-41).
 *     state == NULL  <=>  the reference built with CLOVER_STOCHASTIC_ROUNDING_DISABLED.
 *   - 4-bit values: element 2i in the HIGH nibble, 2i+1 in the LOW nibble of byte i
 *     (include/CloverVector4.h:154-160, 511-514).
 */
#ifndef CLOVER_B200_ORACLE_H
#define CLOVER_B200_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

uint64_t orc_size_pad(uint64_t n);

/* PRNG: include/simdxorshift128plus.h:38-109 */
void orc_xs_init(uint64_t key1, uint64_t key2, uint64_t *state);
void orc_xs_next(uint64_t *state, uint32_t *out8);
void orc_xs_skip(uint64_t *state, uint64_t ncalls);

/* generators: include/CloverVector32.h:712-783 */
void orc_fill_floats(float *x, uint64_t n, float lo, float hi, uint64_t *state);
void orc_fill_integers(float *x, uint64_t n, float lo, float hi, uint64_t *state);

/* CloverVector4: include/CloverVector4.h */
void  orc_v4_quantize(const float *x, uint64_t n, int8_t *values, float *scales, uint64_t *state);
void  orc_v4_restore(const int8_t *values, const float *scales, uint64_t n, float *x);
float orc_v4_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n);
float orc_v4_get(const int8_t *values, const float *scales, uint64_t i);

/* CloverVector8: include/CloverVector8.h */
void  orc_v8_quantize(const float *x, uint64_t n, int8_t *values, float *scales, uint64_t *state);
void  orc_v8_restore(const int8_t *values, const float *scales, uint64_t n, float *x);
float orc_v8_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n);

/* scaleAndAdd (quantized AXPY): r = requantize(u + a * v); r may alias u.
 * include/CloverVector4.h:1222-1478, include/CloverVector8.h:1089-1357 */
void orc_v4_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n,
                          int8_t *r, float *sr, uint64_t *state);
void orc_v8_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n,
                          int8_t *r, float *sr, uint64_t *state);

/* CloverMatrix4: include/CloverMatrix4.h. rows/cols are the PADDED dimensions (multiples of 128). */
void orc_m4_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales, uint64_t *state);
void orc_m4_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32_or_null, uint64_t *state);
/* mixed precision: 4-bit matrix x CloverVector8 -> CloverVector8 (include/CloverMatrix4.h:1093-1441) */
void orc_m4_mvm_v8(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                   const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32_or_null, uint64_t *state);
void orc_m4_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                    const float *x32, float *y32);
/* CloverMatrix8::mvm(V32,V32): include/CloverMatrix8.h:558-661 */
void orc_m8_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                    const float *x32, float *y32);
/* matrix restore: include/CloverMatrix4.h:266-301; 8-bit = get(i, j) element by element (include/CloverMatrix8.h:117-129) */
void orc_m4_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out);
void orc_m8_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out);
/* GEMM definition of this project: C[i][j] = rowView(A,i).dot(rowView(Bt,j)). */
void orc_m4_gemm(const int8_t *av, const float *as, const int8_t *btv, const float *bts,
                 uint64_t K, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, float *c, uint64_t ldc);

/* CloverMatrix8: include/CloverMatrix8.h */
void orc_m8_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales, uint64_t *state);
void orc_m8_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32_or_null, uint64_t *state);


/* transpose: include/CloverMatrix4.h:1549-1663, include/CloverMatrix8.h:1359-1385 */
void orc_m4_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      int8_t *out_values, float *out_scales);
void orc_m8_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      int8_t *out_values, float *out_scales);

/* threshold: include/CloverVector4.h:1913-1973, include/CloverVector8.h:1680-1740 (in place) */
void  orc_v4_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k);
void  orc_v8_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k);
float orc_v4_abs(const int8_t *values, const float *scales, uint64_t i);
float orc_v8_abs(const int8_t *values, const float *scales, uint64_t i);

#ifdef __cplusplus
}
#endif
#endif
