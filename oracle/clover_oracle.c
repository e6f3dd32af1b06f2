/*
 * oracle/clover_oracle.c - plain-C CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see clover_oracle.h). It restates, operation by operation and in
 * the reference's own floating-point ORDER, what the AVX2 code under /root/reference/include
 * computes, so that it can run on a machine where the reference sources are absent (the GPU
 * box). Parity status: PINNED against the compiled reference (oracle/_ref) and tests/golden/.
 *
 * Rules that make the restatement bit-exact (verified, not assumed):
 *   - compiled with -ffp-contract=off; every fused multiply-add is an explicit fmaf()
 *   - 7.0f/max and 127.0f/max are IEEE divisions, never reciprocal-multiplies
 *   - float->int is truncation (cvttps), int->float is round-to-nearest-even (cvtepi32_ps)
 *   - the fp32 dot/mvm accumulation follows the reference's 8-lane FMA chains and its
 *     horizontal-add tree (include/CloverBase.h:149-157)
 */
#include "clover_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define BLOCK 64u

uint64_t orc_size_pad(uint64_t n) { /* include/CloverVector.h:86-89 (CLOVER_VECTOR_SIZE_PAD = 128) */
    return (n % 128u) ? n + 128u - (n % 128u) : n;
}

/* ------------------------------------------------------------------------------------------
 * PRNG. include/simdxorshift128plus.h
 * ---------------------------------------------------------------------------------------- */

/* canonical scalar xorshift128+ step, used only for seeding/jumping (:38-44) */
static void xs_scalar_step(uint64_t *ps0, uint64_t *ps1) {
    uint64_t s1 = *ps0;
    const uint64_t s0 = *ps1;
    *ps0 = s0;
    s1 ^= s1 << 23;
    *ps1 = s1 ^ s0 ^ (s1 >> 18) ^ (s0 >> 5);
}

/* 2^64 jump via the published jump polynomial (:47-62) */
static void xs_scalar_jump(uint64_t in1, uint64_t in2, uint64_t *o1, uint64_t *o2) {
    static const uint64_t JUMP[2] = { 0x8a5cd789635d2dffULL, 0x121fd2155c472f96ULL };
    uint64_t s0 = 0, s1 = 0;
    for (int i = 0; i < 2; ++i)
        for (int b = 0; b < 64; ++b) {
            if (JUMP[i] & (1ULL << b)) { s0 ^= in1; s1 ^= in2; }
            xs_scalar_step(&in1, &in2);
        }
    *o1 = s0; *o2 = s1;
}

/* lane k = lane 0 jumped k * 2^64 canonical steps (:81-92) */
void orc_xs_init(uint64_t key1, uint64_t key2, uint64_t *state) {
    uint64_t *p1 = state, *p2 = state + 4;
    p1[0] = key1; p2[0] = key2;
    for (int k = 1; k < 4; ++k) xs_scalar_jump(p1[k - 1], p2[k - 1], &p1[k], &p2[k]);
}

/* The AVX step AS WRITTEN (:97-109): part1 is overwritten with part2 before use, so each
 * 64-bit lane is the recurrence x' = t ^ x ^ (t >> 18) ^ (x >> 5), t = x ^ (x << 23), and the
 * output is x' + x. out8[2k], out8[2k+1] = low, high half of lane k. */
void orc_xs_next(uint64_t *state, uint32_t *out8) {
    uint64_t *p1 = state, *p2 = state + 4;
    for (int k = 0; k < 4; ++k) {
        const uint64_t s0 = p2[k];
        const uint64_t s1 = s0 ^ (s0 << 23);
        p1[k] = s0;
        p2[k] = (s1 ^ s0 ^ (s1 >> 18)) ^ (s0 >> 5);
        const uint64_t r = p2[k] + s0;
        out8[2 * k] = (uint32_t)r;
        out8[2 * k + 1] = (uint32_t)(r >> 32);
    }
}

void orc_xs_skip(uint64_t *state, uint64_t ncalls) {
    uint32_t sink[8];
    for (uint64_t i = 0; i < ncalls; ++i) orc_xs_next(state, sink);
}

/* ------------------------------------------------------------------------------------------
 * Generators. include/CloverVector32.h:712-783 (same body in CloverMatrix32.h:252-323)
 * ---------------------------------------------------------------------------------------- */
static void fill(float *x, uint64_t n, float lo, float hi, uint64_t *state, int round_int) {
    const float step = (hi - lo) / 2147483648.0f;
    const uint64_t n8 = (n >> 3) << 3;
    uint32_t w[8];
    for (uint64_t i = 0; i < n8; i += 8) {
        orc_xs_next(state, w);
        for (int l = 0; l < 8; ++l) {
            int32_t v = (int32_t)w[l];
            if (v < 0 && v != INT32_MIN) v = -v;          /* _mm256_abs_epi32: abs(INT_MIN) stays INT_MIN */
            float f = fmaf((float)v, step, lo);
            x[i + l] = round_int ? nearbyintf(f) : f;      /* _MM_FROUND_TO_NEAREST_INT */
        }
    }
    for (uint64_t i = n8; i < n; ++i) {                    /* left-overs burn one call each */
        orc_xs_next(state, w);
        int32_t v = (int32_t)w[0];
        if (v < 0 && v != INT32_MIN) v = -v;
        float f = fmaf((float)v, step, lo);
        x[i] = round_int ? nearbyintf(f) : f;
    }
}
void orc_fill_floats(float *x, uint64_t n, float lo, float hi, uint64_t *state) { fill(x, n, lo, hi, state, 0); }
void orc_fill_integers(float *x, uint64_t n, float lo, float hi, uint64_t *state) { fill(x, n, lo, hi, state, 1); }

/* ------------------------------------------------------------------------------------------
 * Shared pieces
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float absf(float f) { return u2f(f2u(f) & 0x7FFFFFFFu); }

/* _mm256_max_ps(a, b) = a > b ? a : b (second operand on ties/NaN) */
static inline float maxps(float a, float b) { return a > b ? a : b; }

/* absmax of 64 floats with the zero guard of include/CloverVector4.h:656-663: a max whose BIT
 * PATTERN is 0 becomes 1.0f (the guard is an integer compare + add of 1.0f, so -0.0 cannot
 * occur: the inputs are already abs()). */
static inline float guard_zero(float m) { return (f2u(m) == 0u) ? 1.0f + m : m; }

/* Stochastic-rounding noise (include/CloverVector4.h:690-734): the call's eight 32-bit words
 * `w`, element slot (g, l): float(int32((w[l] & 0x7F7F7F7F) << 8g)) * 2^-31. */
static inline float noise(const uint32_t *w, int g, int l) {
    const uint32_t m = w[l] & 0x7F7F7F7Fu;
    const int32_t s = (int32_t)(m << (8 * g));
    return (float)s * (1.0f / 2147483648.0f);
}

/* one element: q = sign(x) * trunc(fma(|x|, scale, rnd)); _mm256_sign_epi32 semantics
 * (include/CloverVector4.h:741-772): x bit pattern < 0 -> negate, == 0 -> 0, else keep. */
static inline int32_t quant1(float x, float scale, float rnd) {
    const int32_t q = (int32_t)fmaf(absf(x), scale, rnd);   /* cvttps: truncation */
    const int32_t xi = (int32_t)f2u(x);
    return xi < 0 ? -q : (xi == 0 ? 0 : q);
}

/* Quantize one run of 64 floats with a given (already guarded) max.
 * `transposed` selects the noise mapping: 0 = natural (vector / matrix-tile rows / all 8-bit),
 * 1 = the 4-bit mvm re-quantizer, whose block_values are stored pre-transposed
 * (include/CloverMatrix4.h:806-808, :911, :925-932), i.e. element e takes slot (e%8, e/8). */
static void quant_block(const float *x, float maxv, float qmax, int32_t *q, uint64_t *state, int transposed) {
    const float scale = qmax / maxv;                        /* IEEE divide */
    if (!state) {
        for (int e = 0; e < 64; ++e) q[e] = quant1(x[e], scale, 0.0f);
        return;
    }
    uint32_t w[2][8];
    orc_xs_next(state, w[0]);
    orc_xs_next(state, w[1]);
    for (int e = 0; e < 64; ++e) {
        int c, g, l;
        if (!transposed) { c = e >> 5; g = (e & 31) >> 3; l = e & 7; }
        else             { c = (e & 7) >> 2; g = e & 3; l = e >> 3; }
        q[e] = quant1(x[e], scale, noise(w[c], g, l));
    }
}

static inline void pack4(const int32_t *q, int8_t *dst) {   /* 64 values -> 32 bytes, even element high */
    for (int i = 0; i < 32; ++i)
        dst[i] = (int8_t)(((q[2 * i] & 0xF) << 4) | (q[2 * i + 1] & 0xF));
}
static inline void pack8(const int32_t *q, int8_t *dst) {
    for (int i = 0; i < 64; ++i) dst[i] = (int8_t)q[i];
}
static inline int32_t nib_hi(int8_t b) { return (int32_t)b >> 4; }                    /* sign-extended */
static inline int32_t nib_lo(int8_t b) { return (int32_t)(int8_t)((uint8_t)b << 4) >> 4; }

static inline float block_absmax(const float *x) {
    float m = 0.0f;
    for (int e = 0; e < 64; ++e) m = maxps(absf(x[e]), m);
    return m;
}

/* horizontal add tree of include/CloverBase.h:149-157 */
static inline float hadd8(const float *a) {
    return ((a[4] + a[0]) + (a[6] + a[2])) + ((a[5] + a[1]) + (a[7] + a[3]));
}

/* ------------------------------------------------------------------------------------------
 * CloverVector4
 * ---------------------------------------------------------------------------------------- */
void orc_v4_quantize(const float *x, uint64_t n, int8_t *values, float *scales, uint64_t *state) {
    /* include/CloverVector4.h:605-807 */
    const uint64_t blocks = orc_size_pad(n) / BLOCK;
    int32_t q[64];
    for (uint64_t b = 0; b < blocks; ++b) {
        const float m = guard_zero(block_absmax(x + b * 64));
        scales[b] = m;
        quant_block(x + b * 64, m, 7.0f, q, state, 0);
        pack4(q, values + b * 32);
    }
}

void orc_v4_restore(const int8_t *values, const float *scales, uint64_t n, float *x) {
    /* include/CloverVector4.h:1027-1093 */
    const uint64_t blocks = orc_size_pad(n) / BLOCK;
    for (uint64_t b = 0; b < blocks; ++b) {
        const float s = scales[b] / 7.0f;
        for (int i = 0; i < 32; ++i) {
            x[b * 64 + 2 * i]     = (float)nib_hi(values[b * 32 + i]) * s;
            x[b * 64 + 2 * i + 1] = (float)nib_lo(values[b * 32 + i]) * s;
        }
    }
}

float orc_v4_get(const int8_t *values, const float *scales, uint64_t i) {
    /* include/CloverVector4.h:179-188 */
    const float s = scales[i >> 6] / 7.0f;
    const int8_t b = values[i >> 1];
    return s * (float)((i & 1) ? nib_lo(b) : nib_hi(b));
}

/* scaleAndAdd: r = requantize(u + a * v), block by block (include/CloverVector4.h:1222-1478).
 *   su_ps = su[b] / 7, sv_ps = (sv[b] * a) / 7 (IEEE divides), value = fma(float(qv), sv_ps, float(qu) * su_ps)
 *   absmax -> zero guard -> 7 / max -> truncating quantizer of orc_v4_quantize.
 * The SIMD code peels nibbles by POSITION inside each 32-bit word (slli/srai, :1236-1272): nibble position p of word
 * w is element e with w = e / 8 and p = 2 * ((e / 2) % 4) + (e even); its noise slot is call p / 4, byte p % 4,
 * 32-bit lane w (:1363-1405) - not the slot the quantizer uses. r may alias u. */
void orc_v4_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n,
                          int8_t *r, float *sr, uint64_t *state) {
    const uint64_t blocks = orc_size_pad(n) / BLOCK;
    for (uint64_t b = 0; b < blocks; ++b) {
        const float su_ps = su[b] / 7.0f;
        const float sv_ss = sv[b] * a;
        const float sv_ps = sv_ss / 7.0f;
        float val[64];
        float m = 0.0f;
        for (int i = 0; i < 32; ++i) {
            const int8_t bu = u[b * 32 + i], bv = v[b * 32 + i];
            val[2 * i]     = fmaf((float)nib_hi(bv), sv_ps, (float)nib_hi(bu) * su_ps);
            val[2 * i + 1] = fmaf((float)nib_lo(bv), sv_ps, (float)nib_lo(bu) * su_ps);
        }
        for (int e = 0; e < 64; ++e) m = maxps(absf(val[e]), m);
        m = guard_zero(m);
        sr[b] = m;
        const float scale = 7.0f / m;
        uint32_t w[2][8];
        if (state) { orc_xs_next(state, w[0]); orc_xs_next(state, w[1]); }
        int32_t q[64];
        for (int e = 0; e < 64; ++e) {
            const int p = 2 * ((e >> 1) & 3) + ((e & 1) ? 0 : 1);
            q[e] = quant1(val[e], scale, state ? noise(w[p >> 2], p & 3, e >> 3) : 0.0f);
        }
        pack4(q, r + b * 32);
    }
}

/* 8-bit twin (include/CloverVector8.h:1089-1357): bytes 0..31 of the block use PRNG call 1, bytes 32..63 call 2;
 * byte e of a half sits in 32-bit lane e / 4 at byte position e % 4. */
void orc_v8_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n,
                          int8_t *r, float *sr, uint64_t *state) {
    const uint64_t blocks = orc_size_pad(n) / BLOCK;
    for (uint64_t b = 0; b < blocks; ++b) {
        const float su_ps = su[b] / 127.0f;
        const float sv_ss = sv[b] * a;
        const float sv_ps = sv_ss / 127.0f;
        float val[64];
        float m = 0.0f;
        for (int e = 0; e < 64; ++e) val[e] = fmaf((float)v[b * 64 + e], sv_ps, (float)u[b * 64 + e] * su_ps);
        for (int e = 0; e < 64; ++e) m = maxps(absf(val[e]), m);
        m = guard_zero(m);
        sr[b] = m;
        const float scale = 127.0f / m;
        uint32_t w[2][8];
        if (state) { orc_xs_next(state, w[0]); orc_xs_next(state, w[1]); }
        int32_t q[64];
        for (int e = 0; e < 64; ++e)
            q[e] = quant1(val[e], scale, state ? noise(w[e >> 5], e & 3, (e & 31) >> 2) : 0.0f);
        pack8(q, r + b * 64);
    }
}

/* exact int32 lane sums of one 4-bit block: lane l = bytes 4l..4l+3 = elements 8l..8l+7
 * (include/CloverVector4.h:1134-1181; exact because |q| <= 7 never saturates maddubs) */
static inline void lanes4(const int8_t *u, const int8_t *v, int32_t *lane) {
    for (int l = 0; l < 8; ++l) {
        int32_t s = 0;
        for (int k = 0; k < 4; ++k) {
            const int8_t a = u[4 * l + k], b = v[4 * l + k];
            s += nib_hi(a) * nib_hi(b) + nib_lo(a) * nib_lo(b);
        }
        lane[l] = s;
    }
}

/* The 16 fp32 FMA chains of the SIMD dot (include/CloverVector4.h:1103-1191): even blocks feed
 * accumulator 1, odd blocks accumulator 2, scale = (su * (1/49)) * sv. */
static float dot4_chains(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t blocks) {
    const float rcp49 = 1.0f / 49.0f;
    float acc[2][8];
    int32_t lane[8];
    memset(acc, 0, sizeof acc);
    for (uint64_t b = 0; b < blocks; ++b) {
        const float s = (su[b] * rcp49) * sv[b];
        lanes4(u + b * 32, v + b * 32, lane);
        float *a = acc[b & 1];
        for (int l = 0; l < 8; ++l) a[l] = fmaf(s, (float)lane[l], a[l]);
    }
    float t[8];
    for (int l = 0; l < 8; ++l) t[l] = acc[0][l] + acc[1][l];
    return hadd8(t);
}

float orc_v4_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n) {
    return dot4_chains(u, su, v, sv, orc_size_pad(n) / BLOCK);
}

/* ------------------------------------------------------------------------------------------
 * CloverVector8
 * ---------------------------------------------------------------------------------------- */
void orc_v8_quantize(const float *x, uint64_t n, int8_t *values, float *scales, uint64_t *state) {
    /* include/CloverVector8.h:393-605 */
    const uint64_t blocks = orc_size_pad(n) / BLOCK;
    int32_t q[64];
    for (uint64_t b = 0; b < blocks; ++b) {
        const float m = guard_zero(block_absmax(x + b * 64));
        scales[b] = m;
        quant_block(x + b * 64, m, 127.0f, q, state, 0);
        pack8(q, values + b * 64);
    }
}

void orc_v8_restore(const int8_t *values, const float *scales, uint64_t n, float *x) {
    /* include/CloverVector8.h:835-909 */
    const uint64_t blocks = orc_size_pad(n) / BLOCK;
    for (uint64_t b = 0; b < blocks; ++b) {
        const float s = scales[b] / 127.0f;
        for (int i = 0; i < 64; ++i) x[b * 64 + i] = (float)values[b * 64 + i] * s;
    }
}

/* 8 chains, lane l = bytes 4l..4l+3 and 32+4l..32+4l+3 (include/CloverVector8.h:911-977) */
static float dot8_chains(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t blocks) {
    const float rcp127 = 1.0f / 127.0f;
    float acc[8];
    memset(acc, 0, sizeof acc);
    for (uint64_t b = 0; b < blocks; ++b) {
        const float s = (su[b] * rcp127) * (sv[b] * rcp127);
        const int8_t *pu = u + b * 64, *pv = v + b * 64;
        for (int l = 0; l < 8; ++l) {
            int32_t d = 0;
            for (int k = 0; k < 4; ++k)
                d += (int32_t)pu[4 * l + k] * pv[4 * l + k] + (int32_t)pu[32 + 4 * l + k] * pv[32 + 4 * l + k];
            acc[l] = fmaf(s, (float)d, acc[l]);
        }
    }
    return hadd8(acc);
}

float orc_v8_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n) {
    return dot8_chains(u, su, v, sv, orc_size_pad(n) / BLOCK);
}

/* ------------------------------------------------------------------------------------------
 * Matrices. rows/cols are padded dimensions (include/CloverMatrix.h:48-50).
 * Tiles are visited column-block-major (b_j outer; include/CloverMatrix4.h:524-525), which
 * only matters for the order in which the PRNG stream is consumed (2 calls per tile row).
 * ---------------------------------------------------------------------------------------- */
static void quantize_matrix(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales,
                            uint64_t *state, int bits) {
    const uint64_t hb = cols >> 6, vb = rows >> 6;
    const float qmax = bits == 4 ? 7.0f : 127.0f;
    #pragma omp parallel for collapse(2) schedule(static) if (!state)
    for (uint64_t bj = 0; bj < hb; ++bj)
        for (uint64_t bi = 0; bi < vb; ++bi) {
            const float *t = a + (bi << 6) * cols + (bj << 6);
            float m = 0.0f;
            for (int i = 0; i < 64; ++i) m = maxps(block_absmax(t + (uint64_t)i * cols), m);
            m = guard_zero(m);
            scales[bi * hb + bj] = m;
            int32_t q[64];
            for (int i = 0; i < 64; ++i) {
                quant_block(t + (uint64_t)i * cols, m, qmax, q, state, 0);
                const uint64_t off = ((bi << 6) + (uint64_t)i) * cols + (bj << 6);
                if (bits == 4) pack4(q, values + (off >> 1)); else pack8(q, values + off);
            }
        }
}

void orc_m4_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales, uint64_t *state) {
    quantize_matrix(a, rows, cols, values, scales, state, 4);   /* include/CloverMatrix4.h:512-766 */
}
void orc_m8_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales, uint64_t *state) {
    quantize_matrix(a, rows, cols, values, scales, state, 8);   /* include/CloverMatrix8.h:203-479 */
}

/* mvm(V4,V4): include/CloverMatrix4.h:777-1083. 64 row dots (the inlined SIMD dot against the
 * tile-row's scales), running absmax, then the block-64 re-quantizer with the transposed noise
 * mapping. y32_or_null optionally receives the fp32 intermediates (`block_values`). */
void orc_m4_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32_or_null, uint64_t *state) {
    const uint64_t hb = cols >> 6;
    const uint64_t rb = rows >> 6;
    float *y32 = y32_or_null ? y32_or_null : (float *)malloc(rows * sizeof(float));
    #pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < rows; ++r)
        y32[r] = dot4_chains(values + ((r * cols) >> 1), scales + (r >> 6) * hb, xv, xs, hb);
    for (uint64_t b = 0; b < rb; ++b) {                      /* sequential: consumes the PRNG stream */
        float m = 0.0f;
        for (int i = 0; i < 64; ++i) m = maxps(absf(y32[b * 64 + i]), m);   /* _mm_max_ss(habs, max) */
        m = guard_zero(m);
        ys[b] = m;
        int32_t q[64];
        quant_block(y32 + b * 64, m, 7.0f, q, state, 1);
        pack4(q, yv + b * 32);
    }
    if (!y32_or_null) free(y32);
}

/* mvm(V8,V8): include/CloverMatrix8.h:1002-1298. block_values are in natural order here. */
void orc_m8_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32_or_null, uint64_t *state) {
    const uint64_t hb = cols >> 6;
    const uint64_t rb = rows >> 6;
    float *y32 = y32_or_null ? y32_or_null : (float *)malloc(rows * sizeof(float));
    #pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < rows; ++r)
        y32[r] = dot8_chains(values + r * cols, scales + (r >> 6) * hb, xv, xs, hb);
    for (uint64_t b = 0; b < rb; ++b) {
        float m = 0.0f;
        for (int i = 0; i < 64; ++i) m = maxps(absf(y32[b * 64 + i]), m);
        m = guard_zero(m);
        ys[b] = m;
        int32_t q[64];
        quant_block(y32 + b * 64, m, 127.0f, q, state, 0);
        pack8(q, yv + b * 64);
    }
    if (!y32_or_null) free(y32);
}

/* mvm(V8,V8) on a 4-bit matrix (mixed precision): include/CloverMatrix4.h:1093-1441.
 * Per row 8 fp32 chains; lane l of block b sums elements 4l..4l+3 and 32+4l..32+4l+3 of (4-bit row) x (8-bit x)
 * exactly (maddubs on 16*q never saturates: 2*127*128 < 32768, then >> 4), scale = (su * (1/7)) * (sv * (1/127)),
 * acc_l = fma(scale, float(I_l), acc_l); hadd tree; 8-bit re-quantization of 64 rows, natural noise slots. */
void orc_m4_mvm_v8(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                   const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32_or_null, uint64_t *state) {
    const uint64_t hb = cols >> 6;
    const uint64_t rb = rows >> 6;
    const float rcp7 = 1.0f / 7.0f, rcp127 = 1.0f / 127.0f;
    float *y32 = y32_or_null ? y32_or_null : (float *)malloc(rows * sizeof(float));
    #pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < rows; ++r) {
        const int8_t *u = values + ((r * cols) >> 1);
        const float *su = scales + (r >> 6) * hb;
        float acc[8];
        memset(acc, 0, sizeof acc);
        for (uint64_t b = 0; b < hb; ++b) {
            const float scale = (su[b] * rcp7) * (xs[b] * rcp127);
            for (int l = 0; l < 8; ++l) {
                int32_t sum = 0;
                for (int half = 0; half < 2; ++half)
                    for (int k = 0; k < 4; ++k) {
                        const int e = 32 * half + 4 * l + k;
                        const int8_t byte = u[b * 32 + (e >> 1)];
                        sum += ((e & 1) ? nib_lo(byte) : nib_hi(byte)) * (int32_t)xv[b * 64 + e];
                    }
                acc[l] = fmaf(scale, (float)sum, acc[l]);
            }
        }
        y32[r] = hadd8(acc);
    }
    for (uint64_t b = 0; b < rb; ++b) {
        float m = 0.0f;
        for (int i = 0; i < 64; ++i) m = maxps(absf(y32[b * 64 + i]), m);
        m = guard_zero(m);
        ys[b] = m;
        int32_t q[64];
        quant_block(y32 + b * 64, m, 127.0f, q, state, 0);
        pack8(q, yv + b * 64);
    }
    if (!y32_or_null) free(y32);
}

/* mvm(V32,V32): include/CloverMatrix4.h:1451-1547. 32 chains per row: accumulator k (0..3),
 * lane l, fed per block by element 8k+l and then element 32+8k+l; f = float(q) * (s/7). */
void orc_m4_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                    const float *x32, float *y32) {
    const uint64_t hb = cols >> 6;
    #pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < rows; ++r) {
        const int8_t *u = values + ((r * cols) >> 1);
        const float *su = scales + (r >> 6) * hb;
        float acc[4][8];
        memset(acc, 0, sizeof acc);
        for (uint64_t b = 0; b < hb; ++b) {
            const float s = su[b] / 7.0f;
            const float *x = x32 + b * 64;
            for (int half = 0; half < 2; ++half)
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l < 8; ++l) {
                        const int e = 32 * half + 8 * k + l;
                        const int8_t byte = u[b * 32 + (e >> 1)];
                        const float f = (float)((e & 1) ? nib_lo(byte) : nib_hi(byte)) * s;
                        acc[k][l] = fmaf(x[e], f, acc[k][l]);
                    }
        }
        float t[8];
        for (int l = 0; l < 8; ++l) t[l] = (acc[0][l] + acc[1][l]) + (acc[2][l] + acc[3][l]);
        y32[r] = hadd8(t);
    }
}

/* CloverMatrix8::mvm(V32,V32): include/CloverMatrix8.h:558-661. The same 32 chains as the 4-bit variant (accumulator
 * k = 0..3, AVX lane l; per block element 8k+l, then element 32+8k+l), but the scale is folded into the VECTOR side:
 * s = su[b] / 127.0f, t = x[e] * s (rounded, :632-639), acc = fma(t, float(q[e]), acc) (:641-649). */
void orc_m8_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                    const float *x32, float *y32) {
    const uint64_t hb = cols >> 6;
    #pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < rows; ++r) {
        const int8_t *u = values + r * cols;
        const float *su = scales + (r >> 6) * hb;
        float acc[4][8];
        memset(acc, 0, sizeof acc);
        for (uint64_t b = 0; b < hb; ++b) {
            const float s = su[b] / 127.0f;
            const float *x = x32 + b * 64;
            for (int half = 0; half < 2; ++half)
                for (int k = 0; k < 4; ++k)
                    for (int l = 0; l < 8; ++l) {
                        const int e = 32 * half + 8 * k + l;
                        const float t = x[e] * s;
                        acc[k][l] = fmaf(t, (float)u[b * 64 + e], acc[k][l]);
                    }
        }
        float t[8];
        for (int l = 0; l < 8; ++l) t[l] = (acc[0][l] + acc[1][l]) + (acc[2][l] + acc[3][l]);
        y32[r] = hadd8(t);
    }
}

/* Matrix restore. 4-bit: include/CloverMatrix4.h:266-301 (restore_scalar, the only variant): (s / 7.0f) * float(q).
 * 8-bit: include/CloverMatrix8.h:1300-1309 (restore_scalar) is r[pos] = get(i, j) with get = (s / 127.0f) * float(q)
 * (:117-129); its inner loop increments `i` instead of `j` and never terminates as written, so the restatement is
 * pinned to get(i, j) element by element instead. */
void orc_m4_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out) {
    const uint64_t hb = cols >> 6;
    #pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < rows; ++i)
        for (uint64_t j = 0; j < cols; ++j) {
            const float s = scales[(i >> 6) * hb + (j >> 6)] / 7.0f;
            const int8_t byte = values[(i * cols + j) >> 1];
            out[i * cols + j] = s * (float)((j & 1) ? nib_lo(byte) : nib_hi(byte));
        }
}
void orc_m8_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out) {
    const uint64_t hb = cols >> 6;
    #pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < rows; ++i)
        for (uint64_t j = 0; j < cols; ++j) {
            const float s = scales[(i >> 6) * hb + (j >> 6)] / 127.0f;
            out[i * cols + j] = s * (float)values[i * cols + j];
        }
}

/* GEMM (extension): every C[i][j] is the reference SIMD dot of two row views. */
void orc_m4_gemm(const int8_t *av, const float *as, const int8_t *btv, const float *bts,
                 uint64_t K, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, float *c, uint64_t ldc) {
    const uint64_t kb = K >> 6;
    #pragma omp parallel for schedule(static)
    for (uint64_t i = i0; i < i1; ++i)
        for (uint64_t j = j0; j < j1; ++j)
            c[(i - i0) * ldc + (j - j0)] =
                dot4_chains(av + ((i * K) >> 1), as + (i >> 6) * kb, btv + ((j * K) >> 1), bts + (j >> 6) * kb, kb);
}

/* transpose: include/CloverMatrix4.h:435-502 (scalar), :1549-1663 (AVX2 8x8 nibble blocks),
 * :2508-2640 (parallel), :2649-2802 (faster scalar); include/CloverMatrix8.h:1312-1385. A pure permutation:
 * element (i, j) of the rows x cols source becomes element (j, i) of the cols x rows result, the 64x64-tile scales
 * are transposed likewise (the reference delegates that to ippiTranspose_32f_C1R). Nibble order inside a byte is
 * kept (even column in the high nibble). */
void orc_m4_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      int8_t *out_values, float *out_scales) {
    const uint8_t *u = (const uint8_t *)values;
    uint8_t *v = (uint8_t *)out_values;
    memset(v, 0, rows * cols / 2);
    #pragma omp parallel for schedule(static)
    for (uint64_t j = 0; j < cols; ++j)
        for (uint64_t i = 0; i < rows; ++i) {
            const uint8_t b = u[(i * cols + j) >> 1];
            const uint8_t q = (j & 1) ? (b & 0x0F) : (b >> 4);
            v[(j * rows + i) >> 1] |= (i & 1) ? q : (uint8_t)(q << 4);
        }
    const uint64_t vb = rows >> 6, hb = cols >> 6;
    for (uint64_t bi = 0; bi < vb; ++bi)
        for (uint64_t bj = 0; bj < hb; ++bj) out_scales[bj * vb + bi] = scales[bi * hb + bj];
}

void orc_m8_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      int8_t *out_values, float *out_scales) {
    #pragma omp parallel for schedule(static)
    for (uint64_t j = 0; j < cols; ++j)
        for (uint64_t i = 0; i < rows; ++i) out_values[j * rows + i] = values[i * cols + j];
    const uint64_t vb = rows >> 6, hb = cols >> 6;
    for (uint64_t bi = 0; bi < vb; ++bi)
        for (uint64_t bj = 0; bj < hb; ++bj) out_scales[bj * vb + bi] = scales[bi * hb + bj];
}

/* ---------------------------------------------------------------------------------------------------------
 * threshold: keep the k elements of largest magnitude, zero the rest.
 *   include/CloverVector4.h:1913-1973 (threshold -> threshold_min_heap), include/CloverVector8.h:1680-1740,
 *   heap helpers include/CloverBase.h:205-249. The reference builds a min-heap of the first k elements with
 *   std::make_heap(gt_idx_t) and then replaces the root whenever a later element is STRICTLY larger, re-heapifying
 *   with its own min_heapify. Which of several equal elements survives therefore depends on the heap layout; the
 *   restatement below follows libstdc++'s make_heap (bits/stl_heap.h: __make_heap / __adjust_heap / __push_heap,
 *   GCC 13) and the reference's min_heapify step by step, so it reproduces the reference bit for bit.
 *   getAbs: 4-bit |(scale/7.0f) * (float)q| (:190-203), 8-bit |((float)q * scale) / 127.0f| (CloverVector8.h:141-147).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct { float value; int32_t bits; uint64_t idx; } heap_item;

static int heap_gt(const heap_item *a, const heap_item *b) { return (a->value > b->value) || isnan(a->value); }

static void heap_push(heap_item *first, int64_t hole, int64_t top, heap_item value) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && heap_gt(first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void heap_adjust(heap_item *first, int64_t hole, int64_t len, heap_item value) {
    const int64_t top = hole;
    int64_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (heap_gt(first + child, first + (child - 1))) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    heap_push(first, hole, top, value);
}
static void heap_make(heap_item *first, int64_t len) {
    if (len < 2) return;
    int64_t parent = (len - 2) / 2;
    for (;;) {
        heap_item v = first[parent];
        heap_adjust(first, parent, len, v);
        if (parent == 0) return;
        parent--;
    }
}
static void heap_min_heapify(heap_item *heap, uint32_t pos, uint32_t k) {      /* include/CloverBase.h:226-249 */
    uint32_t smallest = pos;
    for (;;) {
        const uint32_t l = pos * 2 + 1, r = pos * 2 + 2;
        if (l < k && heap[l].value < heap[smallest].value) smallest = l;
        if (r < k && heap[r].value < heap[smallest].value) smallest = r;
        if (smallest == pos) break;
        heap_item t = heap[pos]; heap[pos] = heap[smallest]; heap[smallest] = t;
        pos = smallest;
    }
}

static int32_t v4_getbits(const int8_t *values, uint64_t i) {          /* sign-extended nibble, even index = high nibble */
    const uint8_t b = (uint8_t)values[i >> 1];
    const int32_t nib = (i & 1) ? (b & 0x0F) : (b >> 4);
    return nib >= 8 ? nib - 16 : nib;
}
static void v4_setbits(int8_t *values, uint64_t i, int32_t q) {
    uint8_t b = (uint8_t)values[i >> 1];
    if (i & 1) b = (uint8_t)((b & 0xF0) | (q & 0x0F)); else b = (uint8_t)((b & 0x0F) | ((q & 0x0F) << 4));
    values[i >> 1] = (int8_t)b;
}
float orc_v4_abs(const int8_t *values, const float *scales, uint64_t i) {
    const float scale = scales[i >> 6] / 7.0f;
    return absf(scale * (float)v4_getbits(values, i));
}
float orc_v8_abs(const int8_t *values, const float *scales, uint64_t i) {
    return absf(((float)values[i] * scales[i >> 6]) / 127.0f);
}

static void threshold_impl(int bits, int8_t *values, const float *scales, uint64_t n, uint64_t k) {
    if (k == 0 || k > n) {                                  /* the reference assumes 1 <= k <= n; k = 0 zeroes everything */
        if (k == 0) for (uint64_t i = 0; i < n; ++i) { if (bits == 4) v4_setbits(values, i, 0); else values[i] = 0; }
        return;
    }
    heap_item *heap = (heap_item *)malloc(k * sizeof(heap_item));
    for (uint64_t i = 0; i < k; ++i) {
        heap[i].value = bits == 4 ? orc_v4_abs(values, scales, i) : orc_v8_abs(values, scales, i);
        heap[i].bits = bits == 4 ? v4_getbits(values, i) : values[i];
        heap[i].idx = i;
        if (bits == 4) v4_setbits(values, i, 0); else values[i] = 0;
    }
    heap_make(heap, (int64_t)k);
    for (uint64_t i = k; i < n; ++i) {
        const float value = bits == 4 ? orc_v4_abs(values, scales, i) : orc_v8_abs(values, scales, i);
        if (value > heap[0].value) {
            heap[0].bits = bits == 4 ? v4_getbits(values, i) : values[i];
            heap[0].value = value;
            heap[0].idx = i;
            heap_min_heapify(heap, 0, (uint32_t)k);
        }
        if (bits == 4) v4_setbits(values, i, 0); else values[i] = 0;
    }
    for (uint64_t i = 0; i < k; ++i) {
        if (bits == 4) v4_setbits(values, heap[i].idx, heap[i].bits); else values[heap[i].idx] = (int8_t)heap[i].bits;
    }
    free(heap);
}
void orc_v4_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k) { threshold_impl(4, values, scales, n, k); }
void orc_v8_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k) { threshold_impl(8, values, scales, n, k); }
