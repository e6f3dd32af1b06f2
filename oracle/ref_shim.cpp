/*
 * oracle/ref_shim.cpp - C-callable shim over the UNMODIFIED reference headers.
 *
 * TEST INFRASTRUCTURE ONLY. Compiled by oracle/Makefile from the sources where they lie
 * under /root/reference (never copied into this repo) into oracle/_ref/libclover_ref*.so.
 * Used by tests/ (to pin the C restatement in clover_oracle.c and to generate
 * tests/golden/), by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
 * reference leg. Nothing in the product path (clover_b200/, include/) may link this.
 *
 * Every entry point constructs the reference's own container on caller memory (the
 * reference's borrowing "view" constructors, include/CloverVector4.h:114-119,
 * include/CloverVector8.h:80-85, include/CloverVector32.h:47-51) or, for matrices that
 * have no view constructor, hands out the reference object's own storage.
 *
 * Build flavours (oracle/Makefile):
 *   libclover_ref.so     -DCLOVER_STOCHASTIC_ROUNDING_DISABLED=1   (parity)
 *   libclover_ref_sr.so  stochastic rounding ON (as published, perf baseline + PRNG parity)
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>
#if defined(_OPENMP)
#include <omp.h>
#endif

#include "CloverVector32.h"
#include "CloverVector4.h"
#include "CloverVector8.h"
#include "CloverMatrix32.h"
#include "CloverMatrix4.h"
#include "CloverMatrix8.h"

namespace {

/* The first container ever constructed prints an IPP/MKL banner on stdout
 * (include/CloverBase.h:382-399). Swallow it once so harness output stays parseable. */
struct BannerSilencer {
    BannerSilencer() {
        fflush(stdout);
        int saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (saved >= 0 && nul >= 0) {
            dup2(nul, 1);
            { CloverVector32 warm(128); (void)warm; }
            std::cout.flush();
            fflush(stdout);
            dup2(saved, 1);
        }
        if (nul >= 0) close(nul);
        if (saved >= 0) close(saved);
    }
};
void silence_once() { static BannerSilencer s; (void)s; }

inline __m256i load4(const uint64_t *p) { return _mm256_loadu_si256((const __m256i *)p); }
inline void store4(uint64_t *p, __m256i v) { _mm256_storeu_si256((__m256i *)p, v); }

/* Expose the protected PRNG keys (include/CloverRandom.h:39-41) and storage. */
template <class Base>
struct Keyed : public Base {
    using Base::Base;
    void set_state(const uint64_t *st) { if (st) this->setRandomKeys(load4(st), load4(st + 4)); }
    void get_state(uint64_t *st) { if (st) { store4(st, this->random_key1); store4(st + 4, this->random_key2); } }
    void set_thread_state(const uint64_t *st, int nthreads) {
        if (!st) return;
        for (int t = 0; t < nthreads; ++t) {
            this->random_key1_perthread[t] = load4(st + 8 * t);
            this->random_key2_perthread[t] = load4(st + 8 * t + 4);
        }
    }
};
struct M4 : public Keyed<CloverMatrix4> { using Keyed<CloverMatrix4>::Keyed; int8_t *v() { return values; } float *s() { return scales; } };
struct M8 : public Keyed<CloverMatrix8> { using Keyed<CloverMatrix8>::Keyed; int8_t *v() { return values; } float *s() { return scales; } };
typedef Keyed<CloverVector4> V4;
typedef Keyed<CloverVector8> V8;

} // namespace

extern "C" {

int ref_stochastic_enabled(void) {
#ifdef CLOVER_STOCHASTIC_ROUNDING_DISABLED
    return 0;
#else
    return 1;
#endif
}

int ref_openmp_threads(void) { silence_once(); return CloverBase::get_OpenMP_threads(); }
/* The reference sizes its OpenMP team ONCE, from the first parallel region it runs (include/CloverBase.h:369-380), so
 * this must be the first ref_* call of the process. torchrun exports OMP_NUM_THREADS=1 to every worker: the bench's
 * reference arm overrides that here so that mvm_parallel runs on all host threads. */
void ref_set_openmp_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#endif
}

/* ---- PRNG (include/simdxorshift128plus.h) : state = part1[4] | part2[4] ------------- */
void ref_xs_init(uint64_t key1, uint64_t key2, uint64_t *state) {
    __m256i p1, p2;
    avx_xorshift128plus_init(key1, key2, p1, p2);
    store4(state, p1); store4(state + 4, p2);
}
void ref_xs_next(uint64_t *state, uint32_t *out8) {
    __m256i p1 = load4(state), p2 = load4(state + 4);
    __m256i r = avx_xorshift128plus(p1, p2);
    store4(state, p1); store4(state + 4, p2);
    _mm256_storeu_si256((__m256i *)out8, r);
}

/* ---- fp32 generators (include/CloverVector32.h:712-783) ------------------------------ */
void ref_fill_floats(float *x, uint64_t n, float lo, float hi, uint64_t *state) {
    silence_once();
    CloverVector32 v(n, x);
    __m256i p1 = load4(state), p2 = load4(state + 4);
    v.setRandomFloats(lo, hi, p1, p2);
    store4(state, p1); store4(state + 4, p2);
}
void ref_fill_integers(float *x, uint64_t n, float lo, float hi, uint64_t *state) {
    silence_once();
    CloverVector32 v(n, x);
    __m256i p1 = load4(state), p2 = load4(state + 4);
    v.setRandomInteger(lo, hi, p1, p2);
    store4(state, p1); store4(state + 4, p2);
}

/* ---- CloverVector4 -------------------------------------------------------------------
 * variant: 0 = SIMD (unsuffixed), 1 = _scalar, 2 = _parallel.
 * x must hold size_pad(n) floats (pad zeroed); values n_pad/2 bytes; scales n_pad/64. */
void ref_v4_quantize(const float *x, uint64_t n, int8_t *values, float *scales, uint64_t *state, int variant) {
    silence_once();
    CloverVector32 in(n, (float *)x);
    V4 q(n, values, scales);
    q.set_state(state);
    if (variant == 1) q.quantize_scalar(in); else if (variant == 2) q.quantize_parallel(in); else q.quantize(in);
    q.get_state(state);
}
void ref_v4_restore(const int8_t *values, const float *scales, uint64_t n, float *x, int variant) {
    silence_once();
    CloverVector32 out(n, x);
    CloverVector4 q(n, (int8_t *)values, (float *)scales);
    if (variant == 1) q.restore_scalar(out); else q.restore(out);
}
float ref_v4_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n, int variant) {
    silence_once();
    CloverVector4 a(n, (int8_t *)u, (float *)su);
    CloverVector4 b(n, (int8_t *)v, (float *)sv);
    if (variant == 1) return a.dot_scalar(b);
    if (variant == 2) return a.dot_parallel(b);
    return a.dot(b);
}
float ref_v4_get(const int8_t *values, const float *scales, uint64_t n, uint64_t i) {
    silence_once();
    CloverVector4 q(n, (int8_t *)values, (float *)scales);
    return q.get(i);
}
uint64_t ref_v4_bytes(uint64_t n) { silence_once(); CloverVector4 q(n, nullptr, nullptr); return q.getBytes(); }
uint64_t ref_v_size_pad(uint64_t n) { silence_once(); CloverVector4 q(n, nullptr, nullptr); return q.size_pad(); }

/* ---- CloverVector8 ------------------------------------------------------------------- */
void ref_v8_quantize(const float *x, uint64_t n, int8_t *values, float *scales, uint64_t *state, int variant) {
    silence_once();
    CloverVector32 in(n, (float *)x);
    V8 q(n, values, scales);
    q.set_state(state);
    if (variant == 1) q.quantize_scalar(in); else if (variant == 2) q.quantize_parallel(in); else q.quantize(in);
    q.get_state(state);
}
void ref_v8_restore(const int8_t *values, const float *scales, uint64_t n, float *x, int variant) {
    silence_once();
    CloverVector32 out(n, x);
    CloverVector8 q(n, (int8_t *)values, (float *)scales);
    if (variant == 1) q.restore_scalar(out); else q.restore(out);
}
float ref_v8_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n, int variant) {
    silence_once();
    CloverVector8 a(n, (int8_t *)u, (float *)su);
    CloverVector8 b(n, (int8_t *)v, (float *)sv);
    if (variant == 1) return a.dot_scalar(b);
    if (variant == 2) return a.dot_parallel(b);
    return a.dot(b);
}

/* ---- scaleAndAdd: r = u + a * v, re-quantized. variant: 0 = SIMD, 1 = _scalar, 2 = _parallel ---------------- */
void ref_v4_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n,
                          int8_t *r, float *sr, uint64_t *state, int variant) {
    silence_once();
    V4 x(n, (int8_t *)u, (float *)su);
    CloverVector4 y(n, (int8_t *)v, (float *)sv);
    CloverVector4 out(n, r, sr);
    x.set_state(state);
    if (variant == 1) x.scaleAndAdd_scalar(y, a, out); else if (variant == 2) x.scaleAndAdd_parallel(y, a, out); else x.scaleAndAdd(y, a, out);
    x.get_state(state);
}
void ref_v8_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n,
                          int8_t *r, float *sr, uint64_t *state, int variant) {
    silence_once();
    V8 x(n, (int8_t *)u, (float *)su);
    CloverVector8 y(n, (int8_t *)v, (float *)sv);
    CloverVector8 out(n, r, sr);
    x.set_state(state);
    if (variant == 1) x.scaleAndAdd_scalar(y, a, out); else if (variant == 2) x.scaleAndAdd_parallel(y, a, out); else x.scaleAndAdd(y, a, out);
    x.get_state(state);
}

/* threshold in place (include/CloverVector4.h:1913, CloverVector8.h:1680); variant 2 = threshold_parallel */
void ref_v4_threshold(int8_t *values, float *scales, uint64_t n, uint64_t k, int variant) {
    silence_once();
    CloverVector4 q(n, values, scales);
    if (variant == 2) q.threshold_parallel(k); else q.threshold(k);
}
void ref_v8_threshold(int8_t *values, float *scales, uint64_t n, uint64_t k, int variant) {
    silence_once();
    CloverVector8 q(n, values, scales);
    if (variant == 2) q.threshold_parallel(k); else q.threshold(k);
}

/* ---- CloverMatrix32 (input container) ------------------------------------------------ */
void *ref_m32_create(uint64_t rows, uint64_t cols) { silence_once(); return new CloverMatrix32(rows, cols); }
void ref_m32_destroy(void *h) { delete (CloverMatrix32 *)h; }
float *ref_m32_data(void *h) { return ((CloverMatrix32 *)h)->getData(); }
uint64_t ref_m32_rows(void *h) { return ((CloverMatrix32 *)h)->getRows(); }
uint64_t ref_m32_cols(void *h) { return ((CloverMatrix32 *)h)->getCols(); }
void ref_m32_fill_floats(void *h, float lo, float hi, uint64_t *state) {
    __m256i p1 = load4(state), p2 = load4(state + 4);
    ((CloverMatrix32 *)h)->setRandomFloats(lo, hi, p1, p2);
    store4(state, p1); store4(state + 4, p2);
}
void ref_m32_fill_integers(void *h, float lo, float hi, uint64_t *state) {
    __m256i p1 = load4(state), p2 = load4(state + 4);
    ((CloverMatrix32 *)h)->setRandomInteger(lo, hi, p1, p2);
    store4(state, p1); store4(state + 4, p2);
}

/* ---- CloverMatrix4 ------------------------------------------------------------------- */
void *ref_m4_create(uint64_t rows, uint64_t cols) { silence_once(); return new M4(rows, cols); }
void ref_m4_destroy(void *h) { delete (M4 *)h; }
int8_t *ref_m4_values(void *h) { return ((M4 *)h)->v(); }
float *ref_m4_scales(void *h) { return ((M4 *)h)->s(); }
uint64_t ref_m4_rows(void *h) { return ((M4 *)h)->getRows(); }
uint64_t ref_m4_cols(void *h) { return ((M4 *)h)->getCols(); }
uint64_t ref_m4_bytes(void *h) { return ((M4 *)h)->getBytes(); }
float ref_m4_get(void *h, uint64_t i, uint64_t j) { return ((M4 *)h)->get(i, j); }
void ref_m4_quantize(void *h, void *m32, uint64_t *state, int variant) {
    M4 *m = (M4 *)h;
    m->set_state(state);
    if (variant == 1) m->quantize_scalar(*(CloverMatrix32 *)m32); else m->quantize(*(CloverMatrix32 *)m32);
    m->get_state(state);
}
/* x: V4 of length cols, y: V4 of length rows (include/CloverMatrix4.h:777, :311, :1681). */
void ref_m4_mvm(void *h, const int8_t *xv, const float *xs, int8_t *yv, float *ys, uint64_t *state, int variant) {
    M4 *m = (M4 *)h;
    CloverVector4 x(m->getCols(), (int8_t *)xv, (float *)xs);
    CloverVector4 y(m->getRows(), yv, ys);
    m->set_state(state);
    if (variant == 1) m->mvm_scalar(x, y); else if (variant == 2) m->mvm_parallel(x, y); else m->mvm(x, y);
    m->get_state(state);
}
/* mixed precision: CloverVector8 x / y on the 4-bit matrix (include/CloverMatrix4.h:1093; _parallel :2017). */
void ref_m4_mvm_v8(void *h, const int8_t *xv, const float *xs, int8_t *yv, float *ys, uint64_t *state, int variant) {
    M4 *m = (M4 *)h;
    CloverVector8 x(m->getCols(), (int8_t *)xv, (float *)xs);
    CloverVector8 y(m->getRows(), yv, ys);
    m->set_state(state);
    if (variant == 2) m->mvm_parallel(x, y); else m->mvm(x, y);
    m->get_state(state);
}
/* fp32 x / fp32 y (include/CloverMatrix4.h:1451, :423, :2397). */
void ref_m4_mvm_f32(void *h, const float *x32, float *y32, int variant) {
    M4 *m = (M4 *)h;
    CloverVector32 x(m->getCols(), (float *)x32);
    CloverVector32 y(m->getRows(), y32);
    if (variant == 1) m->mvm_scalar(x, y); else if (variant == 2) m->mvm_parallel(x, y); else m->mvm(x, y);
}
/* The GEMM definition adopted by this project: C[i][j] = rowView(A,i).dot(rowView(Bt,j)),
 * built with exactly the view construction the reference's own mvm_scalar uses (include/CloverMatrix4.h:338-342). */
void ref_m4_gemm_rows(void *ha, void *hbt, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, float *c, uint64_t ldc) {
    M4 *a = (M4 *)ha, *bt = (M4 *)hbt;
    const uint64_t K = a->getCols();
    const uint64_t kb = K >> 6;
    #pragma omp parallel for schedule(static)
    for (uint64_t i = i0; i < i1; ++i) {
        CloverVector4 ra(K, a->v() + ((i * K) >> 1), a->s() + (i >> 6) * kb);
        for (uint64_t j = j0; j < j1; ++j) {
            CloverVector4 rb(K, bt->v() + ((j * K) >> 1), bt->s() + (j >> 6) * kb);
            c[(i - i0) * ldc + (j - j0)] = ra.dot(rb);
        }
    }
}

/* transpose into another reference matrix (include/CloverMatrix4.h:1549 SIMD, :435 scalar, :2508 parallel, :2649 faster scalar) */
void ref_m4_transpose(void *h, void *hout, int variant) {
    M4 *m = (M4 *)h, *o = (M4 *)hout;
    if (variant == 1) m->transpose_scalar(*o); else if (variant == 2) m->transpose_parallel(*o);
    else if (variant == 3) m->transpose_scalar_faster(*o); else m->transpose(*o);
}

/* ---- CloverMatrix8 ------------------------------------------------------------------- */
void *ref_m8_create(uint64_t rows, uint64_t cols) { silence_once(); return new M8(rows, cols); }
void ref_m8_destroy(void *h) { delete (M8 *)h; }
int8_t *ref_m8_values(void *h) { return ((M8 *)h)->v(); }
float *ref_m8_scales(void *h) { return ((M8 *)h)->s(); }
uint64_t ref_m8_rows(void *h) { return ((M8 *)h)->getRows(); }
uint64_t ref_m8_cols(void *h) { return ((M8 *)h)->getCols(); }
uint64_t ref_m8_bytes(void *h) { return ((M8 *)h)->getBytes(); }
void ref_m8_quantize(void *h, void *m32, uint64_t *state, int variant) {
    M8 *m = (M8 *)h;
    m->set_state(state);
    if (variant == 1) m->quantize_scalar(*(CloverMatrix32 *)m32); else m->quantize(*(CloverMatrix32 *)m32);
    m->get_state(state);
}
void ref_m8_mvm(void *h, const int8_t *xv, const float *xs, int8_t *yv, float *ys, uint64_t *state, int variant) {
    M8 *m = (M8 *)h;
    CloverVector8 x(m->getCols(), (int8_t *)xv, (float *)xs);
    CloverVector8 y(m->getRows(), yv, ys);
    m->set_state(state);
    if (variant == 1) m->mvm_scalar(x, y); else if (variant == 2) m->mvm_parallel(x, y); else m->mvm(x, y);
    m->get_state(state);
}

/* fp32 x / fp32 y on the 8-bit matrix (include/CloverMatrix8.h:558 SIMD, :546 scalar with a double accumulator). */
void ref_m8_mvm_f32(void *h, const float *x32, float *y32, int variant) {
    M8 *m = (M8 *)h;
    CloverVector32 x(m->getCols(), (float *)x32);
    CloverVector32 y(m->getRows(), y32);
    if (variant == 1) m->mvm_scalar(x, y); else m->mvm(x, y);
}
/* matrix restore: the 4-bit class has restore_scalar only (include/CloverMatrix4.h:266); the 8-bit restore_scalar
 * (include/CloverMatrix8.h:1300) does not terminate as written, so the 8-bit matrix is read through get(i, j) (:117). */
void ref_m4_restore(void *h, float *out) {
    M4 *m = (M4 *)h;
    CloverMatrix32 *o = new CloverMatrix32(m->getRows(), m->getCols());
    m->restore_scalar(*o);
    memcpy(out, o->getData(), (size_t)m->getRows() * m->getCols() * sizeof(float));
    delete o;
}
float ref_m8_get(void *h, uint64_t i, uint64_t j) { return ((M8 *)h)->get(i, j); }
void ref_m8_restore_by_get(void *h, float *out) {
    M8 *m = (M8 *)h;
    const uint64_t R = m->getRows(), Cc = m->getCols();
    for (uint64_t i = 0; i < R; ++i)
        for (uint64_t j = 0; j < Cc; ++j) out[i * Cc + j] = m->get(i, j);
}

/* include/CloverMatrix8.h:1359 (IPP), :1312 scalar, :1338 parallel */
void ref_m8_transpose(void *h, void *hout, int variant) {
    M8 *m = (M8 *)h, *o = (M8 *)hout;
    if (variant == 1) m->transpose_scalar(*o); else if (variant == 2) m->transpose_parallel(*o); else m->transpose(*o);
}

} // extern "C"
