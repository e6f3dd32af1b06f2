// axpy_kernels.cu - CloverVector4 / CloverVector8 :: scaleAndAdd, the quantized AXPY  r = requantize(u + a * v).
//
// Reference: include/CloverVector4.h:1222-1478, include/CloverVector8.h:1089-1357 (the step either
// side of mvm in the IHT / gradient-descent loops, test/performance/01_measure.h:940-944). Per block of 64:
//     su_ps = su[b] / 7,  sv_ps = (sv[b] * a) / 7                     (IEEE divides; 127 for 8-bit)
//     value = fma(float(qv), sv_ps, float(qu) * su_ps)
//     m = absmax(value), zero guard, scale = 7 / m, q = sign * trunc(fma(|value|, scale, noise)), pack; sr[b] = m
// One HBM pass (1.6875 B/element for 4-bit, 3.1875 for 8-bit, the reference's getBytes() model). One thread owns one
// block at a time (grid-stride), so r may alias u exactly like the reference's two-argument form.
//
// Stochastic rounding: the SIMD code peels nibbles by POSITION inside each 32-bit word, so element e of a 4-bit
// block takes noise slot (call p / 4, byte p % 4, lane e / 8) with p = 2 * ((e / 2) % 4) + (e even); an 8-bit element
// takes (call e / 32, byte e % 4, lane (e % 32) / 4). The thread jumps the object's XORShift stream to 2 * block.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

__device__ __forceinline__ uint32_t prmt_b(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// byte j of `biased` (a value 0..255 that encodes x + bias) -> float(x), without I2F: splice the byte under the
// bits of 1.5 * 2^23 and subtract 12582912 + bias (exact).
template <int J>
__device__ __forceinline__ float byte_to_float(uint32_t biased, float magic_plus_bias) {
    const uint32_t bits = prmt_b(biased, 0x4B400000u, 0x7650u | J);        // {magic[3], magic[2], magic[1], biased[J]}
    return __fadd_rn(__uint_as_float(bits), -magic_plus_bias);
}

// ---- packed (f32x2) arithmetic of the rounding-disabled path -----------------------------------------------------
// SASS PRMT / LOP3 take one immediate: the magic bits and the masks live in registers (built from a kernel argument
// that is always 0, so that ptxas cannot fold them back into immediates - see mvm_f32_kernels.cu).
struct AxpyConsts { uint32_t magic, m0f, c08, c80; };
__device__ __forceinline__ uint32_t and_xor_r(uint32_t a, uint32_t m, uint32_t c) {      // (a & m) ^ c in one LOP3
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
    return d;
}
// val2 = fma(float(qv2), sv, float(qu2) * su) for the two elements whose biased codes sit in byte P of (ua, ub) / (va, vb):
// value = fma(qv, sv_ps, qu * su_ps) (include/CloverVector4.h:1300-1340), both halves rounded like the scalar instructions
template <int PA, int PB>
__device__ __forceinline__ uint64_t axpy_pair(uint32_t ua, uint32_t ub, uint32_t va, uint32_t vb, uint32_t magic, uint64_t neg,
                                              uint64_t su2, uint64_t sv2) {
    uint64_t r;
    asm("{\n\t.reg .b32 a, b, c, d;\n\t.reg .b64 mu, mv;\n\t"
        "prmt.b32 a, %1, %5, %9;\n\t"
        "prmt.b32 b, %2, %5, %10;\n\t"
        "prmt.b32 c, %3, %5, %9;\n\t"
        "prmt.b32 d, %4, %5, %10;\n\t"
        "mov.b64 mu, {a, b};\n\t"
        "mov.b64 mv, {c, d};\n\t"
        "add.rn.f32x2 mu, mu, %6;\n\t"
        "add.rn.f32x2 mv, mv, %6;\n\t"
        "mul.rn.f32x2 mu, mu, %7;\n\t"
        "fma.rn.f32x2 %0, mv, %8, mu;\n\t}"
        : "=l"(r) : "r"(ua), "r"(ub), "r"(va), "r"(vb), "r"(magic), "l"(neg), "l"(su2), "l"(sv2), "n"(0x7650 + PA), "n"(0x7650 + PB));
    return r;
}
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ float max3_abs(float m, float a, float b) {                    // one FMNMX3 on sm_100
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return d;
}
// d = (c << 8) | (nibble(a) << 4) | nibble(b): one instruction per byte of the reference's nibble layout
__device__ __forceinline__ uint32_t pack_s4(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.s4.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// d = (c << 16) | (byte(a) << 8) | byte(b)
__device__ __forceinline__ uint32_t pack_s8(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---------------------------------------------------------------------------------------------
// Keyed (stochastic-rounding) kernel: FOUR threads per block of 64, 16 contiguous elements each - the layout that took the
// quantizer from 73 % to 93 % of the HBM roofline (vector_kernels.cu, k_vquantize4t). A thread loads its 8 (4-bit) or
// 16 (8-bit) bytes of u and v with one vector load each, the block absmax is two xor-shuffles away, and its 16
// roundings pack into exactly the bytes it loaded - so r may still alias u: every byte is read and written by the same
// thread, and sr[b] is written (by thread 0 of the group) only after the shuffles, which all four threads reach after
// their read of su[b] has returned. ~45 registers instead of 82-115: full occupancy.
// Stochastic mode: a group walks R consecutive blocks and all four threads step the same XORShift state.
// ---------------------------------------------------------------------------------------------
template <int BITS, bool STOCH>
__global__ void __launch_bounds__(256)
k_vscale_add4t(const uint32_t *u, const float *su, const uint32_t *__restrict__ v, const float *__restrict__ sv, float a,
               uint64_t nblocks, uint64_t R, uint32_t *r, float *sr, Key4 key, const uint64_t *__restrict__ tables, uint32_t zero) {
    constexpr float kQmax = BITS == 4 ? 7.0f : 127.0f;
    const AxpyConsts kc = {0x4B400000u | zero, 0x0F0F0F0Fu | zero, 0x08080808u | zero, 0x80808080u | zero};
    constexpr int kW = BITS == 4 ? 2 : 4;                          // 32-bit words per thread
    const int s = threadIdx.x & 3;
    const uint64_t group = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 2;
    const uint64_t ngroups = ((uint64_t)gridDim.x * 256) >> 2;
    uint64_t blk = STOCH ? group * R : group;
    const uint64_t step = STOCH ? 1 : ngroups;
    const uint64_t end = STOCH ? min(nblocks, (group + 1) * R) : nblocks;
    uint64_t lanes[4] = {0, 0, 0, 0};
    if (STOCH && blk < end) {
#pragma unroll
        for (int k = 0; k < 4; ++k) lanes[k] = xs_jump(tables, key.x[k], 2 * blk);
    }
    // Operands of the NEXT block are requested before the current one is computed: with a single block in flight per
    // thread the kernel ran at DRAM latency, not bandwidth (56 us at n = 2^26 whatever the instruction count - r02h).
    // In-place use stays safe: a thread only ever reads and writes its own bytes, and the next block's are not written yet.
    uint32_t wu[kW], wv[kW], nwu[kW], nwv[kW];
    float su_raw = 0.f, sv_raw = 0.f, nsu_raw = 0.f, nsv_raw = 0.f;
    auto fetch = [&](uint64_t b, uint32_t *pu, uint32_t *pv, float &psu, float &psv) {
#pragma unroll
        for (int i = 0; i < kW; ++i) pu[i] = pv[i] = 0u;
        psu = psv = 0.f;
        if (b < end) {
            psu = su[b];
            psv = sv[b];
            if (BITS == 4) {
                const uint2 x = *reinterpret_cast<const uint2 *>(u + b * 8 + 2 * s), y = *reinterpret_cast<const uint2 *>(v + b * 8 + 2 * s);
                pu[0] = x.x; pu[1] = x.y; pv[0] = y.x; pv[1] = y.y;
            } else {
                const uint4 x = *reinterpret_cast<const uint4 *>(u + b * 16 + 4 * s), y = *reinterpret_cast<const uint4 *>(v + b * 16 + 4 * s);
                pu[0] = x.x; pu[1] = x.y; pu[2] = x.z; pu[3] = x.w; pv[0] = y.x; pv[1] = y.y; pv[2] = y.z; pv[3] = y.w;
            }
        }
    };
    auto advance = [&]() {
#pragma unroll
        for (int i = 0; i < kW; ++i) { wu[i] = nwu[i]; wv[i] = nwv[i]; }
        su_raw = nsu_raw; sv_raw = nsv_raw;
    };
    fetch(blk, wu, wv, su_raw, sv_raw);
    for (;; blk += step, advance()) {
        const bool live = blk < end;
        if (!__any_sync(0xFFFFFFFFu, live)) break;
        fetch(blk + step, nwu, nwv, nsu_raw, nsv_raw);
        float su_ps = 0.f, sv_ps = 0.f;
        if (live) {
            su_ps = __fdiv_rn(su_raw, kQmax);
            sv_ps = __fdiv_rn(__fmul_rn(sv_raw, a), kQmax);
        }
        float val[16];
        if (BITS == 4) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const uint32_t bu = wu[w] ^ 0x88888888u, bv = wv[w] ^ 0x88888888u;
                const uint32_t uh = (bu >> 4) & 0x0F0F0F0Fu, ul = bu & 0x0F0F0F0Fu;
                const uint32_t vh = (bv >> 4) & 0x0F0F0F0Fu, vl = bv & 0x0F0F0F0Fu;
#define CLOVER_AXPY4T(J)                                                                                             \
                val[8 * w + 2 * J]     = __fmaf_rn(byte_to_float<J>(vh, 12582920.0f), sv_ps,                          \
                                                   __fmul_rn(byte_to_float<J>(uh, 12582920.0f), su_ps));             \
                val[8 * w + 2 * J + 1] = __fmaf_rn(byte_to_float<J>(vl, 12582920.0f), sv_ps,                          \
                                                   __fmul_rn(byte_to_float<J>(ul, 12582920.0f), su_ps));
                CLOVER_AXPY4T(0) CLOVER_AXPY4T(1) CLOVER_AXPY4T(2) CLOVER_AXPY4T(3)
#undef CLOVER_AXPY4T
            }
        } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t bu = wu[w] ^ 0x80808080u, bv = wv[w] ^ 0x80808080u;
#define CLOVER_AXPY8T(J)                                                                                             \
                val[4 * w + J] = __fmaf_rn(byte_to_float<J>(bv, 12583040.0f), sv_ps,                                  \
                                           __fmul_rn(byte_to_float<J>(bu, 12583040.0f), su_ps));
                CLOVER_AXPY8T(0) CLOVER_AXPY8T(1) CLOVER_AXPY8T(2) CLOVER_AXPY8T(3)
#undef CLOVER_AXPY8T
            }
        }
        float m = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) m = fmaxf(m, fabsf(val[e]));
        if (!live) m = 0.f;
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
        m = guard_zero(m);
        const float scale = quant_scale(kQmax, m);
        // noise words: 4-bit - word W of the block takes PRNG word W of call 0 (nibble positions 0..3) and of call 1
        // (positions 4..7); 8-bit - element e takes (call e / 32, word (e % 32) / 4, byte e % 4)
        uint32_t nw[2][kW];
        if (STOCH) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t o = xs_next(lanes[k]);
                    const uint32_t lo = (uint32_t)o, hi = (uint32_t)(o >> 32);
                    if (BITS == 4) {                      // this thread's words 2s, 2s+1 = PRNG words of 64-bit lane k = s
                        if (k == s) { nw[c][0] = lo; nw[c][1] = hi; }
                    } else if (c == (s >> 1)) {           // elements 16s..16s+15: words (4s + j) % 8, j = 0..3 = lanes 2(s&1), 2(s&1)+1
                        if (k == 2 * (s & 1)) { nw[0][0] = lo; nw[0][1] = hi; }
                        if (k == 2 * (s & 1) + 1) { nw[0][2] = lo; nw[0][3] = hi; }
                    }
                }
        }
        if (!live) continue;
        if (s == 0) sr[blk] = m;
        if (BITS == 4) {
            uint32_t out[2];
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                int q[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pnib = 2 * ((i >> 1) & 3) + ((i & 1) ? 0 : 1);                 // nibble position of element 8w+i
                    const float rnd = STOCH ? noise_from_word(nw[pnib >> 2][w], pnib & 3) : 0.f;
                    q[i] = quant_one(val[8 * w + i], scale, rnd);
                }
                out[w] = pack8_nibbles(q);
            }
            *reinterpret_cast<uint2 *>(r + blk * 8 + 2 * s) = make_uint2(out[0], out[1]);
        } else {
            uint32_t out[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                int q[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float rnd = STOCH ? noise_from_word(nw[0][w], i) : 0.f;
                    q[i] = quant_one(val[4 * w + i], scale, rnd);
                }
                out[w] = pack4_bytes(q);
            }
            *reinterpret_cast<uint4 *>(r + blk * 16 + 4 * s) = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Rounding-disabled kernel, TPB threads per block of 64 (EPT = 64 / TPB elements each), packed arithmetic throughout.
// Everything a block costs ONCE - two operand-scale divides, the 7 / max divide, address arithmetic, the loop - is paid
// per thread, so fewer threads per block means fewer instructions per element: ncu counted 15.3 instructions per
// element with four threads per block against 7.3 in the arithmetic itself (profiles/r02p). r may alias u (every byte
// is read and written by the same thread; the next block's operands are prefetched before the current block is stored).
// ---------------------------------------------------------------------------------------------
template <int BITS, int TPB>
__global__ void __launch_bounds__(256)
k_vscale_add_pk(const uint32_t *u, const float *su, const uint32_t *__restrict__ v, const float *__restrict__ sv, float a,
                uint64_t nblocks, uint32_t *r, float *sr, uint32_t zero) {
    constexpr float kQmax = BITS == 4 ? 7.0f : 127.0f;
    constexpr int EPT = 64 / TPB;                                  // elements per thread
    constexpr int kW = BITS == 4 ? EPT / 8 : EPT / 4;              // 32-bit words per thread
    constexpr int kBW = BITS == 4 ? 8 : 16;                        // words per block
    const AxpyConsts kc = {0x4B400000u | zero, 0x0F0F0F0Fu | zero, 0x08080808u | zero, 0x80808080u | zero};
    const int s = threadIdx.x & (TPB - 1);
    const uint64_t group = ((uint64_t)blockIdx.x * 256 + threadIdx.x) / TPB, ngroups = ((uint64_t)gridDim.x * 256) / TPB;
    uint32_t wu[kW], wv[kW], nwu[kW], nwv[kW];
    float su_raw, sv_raw, nsu_raw, nsv_raw;
    auto fetch = [&](uint64_t b, uint32_t *pu, uint32_t *pv, float &psu, float &psv) {
#pragma unroll
        for (int i = 0; i < kW; ++i) pu[i] = pv[i] = 0u;
        psu = psv = 0.f;
        if (b < nblocks) {
            psu = su[b];
            psv = sv[b];
            const uint32_t *gu = u + b * kBW + s * kW, *gv = v + b * kBW + s * kW;
            if (kW == 2) {
                const uint2 x = *reinterpret_cast<const uint2 *>(gu), y = *reinterpret_cast<const uint2 *>(gv);
                pu[0] = x.x; pu[1] = x.y; pv[0] = y.x; pv[1] = y.y;
            } else {
#pragma unroll
                for (int i = 0; i < kW / 4; ++i) {
                    const uint4 x = reinterpret_cast<const uint4 *>(gu)[i], y = reinterpret_cast<const uint4 *>(gv)[i];
                    pu[4 * i] = x.x; pu[4 * i + 1] = x.y; pu[4 * i + 2] = x.z; pu[4 * i + 3] = x.w;
                    pv[4 * i] = y.x; pv[4 * i + 1] = y.y; pv[4 * i + 2] = y.z; pv[4 * i + 3] = y.w;
                }
            }
        }
    };
    uint64_t blk = group;
    fetch(blk, wu, wv, su_raw, sv_raw);
    for (;; blk += ngroups) {
        const bool live = blk < nblocks;
        if (!__any_sync(0xFFFFFFFFu, live)) break;
        fetch(blk + ngroups, nwu, nwv, nsu_raw, nsv_raw);
        const float su_ps = __fdiv_rn(su_raw, kQmax), sv_ps = __fdiv_rn(__fmul_rn(sv_raw, a), kQmax);
        const uint64_t su2 = f2_pack(su_ps, su_ps), sv2 = f2_pack(sv_ps, sv_ps);
        uint64_t val2[EPT / 2];
        if (BITS == 4) {
            const uint64_t neg = f2_pack(-12582920.0f, -12582920.0f);
#pragma unroll
            for (int w = 0; w < kW; ++w) {                            // byte p of a word: element 2p in the HIGH nibble, 2p+1 in the low one
                const uint32_t ul = and_xor_r(wu[w], kc.m0f, kc.c08), uh = and_xor_r(wu[w] >> 4, kc.m0f, kc.c08);
                const uint32_t vl = and_xor_r(wv[w], kc.m0f, kc.c08), vh = and_xor_r(wv[w] >> 4, kc.m0f, kc.c08);
                val2[4 * w + 0] = axpy_pair<0, 0>(uh, ul, vh, vl, kc.magic, neg, su2, sv2);
                val2[4 * w + 1] = axpy_pair<1, 1>(uh, ul, vh, vl, kc.magic, neg, su2, sv2);
                val2[4 * w + 2] = axpy_pair<2, 2>(uh, ul, vh, vl, kc.magic, neg, su2, sv2);
                val2[4 * w + 3] = axpy_pair<3, 3>(uh, ul, vh, vl, kc.magic, neg, su2, sv2);
            }
        } else {
            const uint64_t neg = f2_pack(-12583040.0f, -12583040.0f);
#pragma unroll
            for (int w = 0; w < kW; ++w) {
                const uint32_t bu = wu[w] ^ kc.c80, bv = wv[w] ^ kc.c80;              // q + 128 per byte
                val2[2 * w + 0] = axpy_pair<0, 1>(bu, bu, bv, bv, kc.magic, neg, su2, sv2);
                val2[2 * w + 1] = axpy_pair<2, 3>(bu, bu, bv, bv, kc.magic, neg, su2, sv2);
            }
        }
        float m = 0.f;
#pragma unroll
        for (int e = 0; e < EPT / 2; ++e) m = max3_abs(m, f2_lo(val2[e]), f2_hi(val2[e]));
        if (!live) m = 0.f;
#pragma unroll
        for (int o = 1; o < TPB; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        m = guard_zero(m);
        if (live) {
            // q = sign(val) * trunc(fma(|val|, scale, 0)) = trunc(val * scale): rounding to nearest is sign-symmetric, truncation is odd
            const float scale = quant_scale(kQmax, m);
            const uint64_t sc2 = f2_pack(scale, scale);
            if (s == 0) sr[blk] = m;
            int q[EPT];
#pragma unroll
            for (int e = 0; e < EPT / 2; ++e) {
                uint64_t p;
                asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(val2[e]), "l"(sc2));
                q[2 * e] = __float2int_rz(f2_lo(p));
                q[2 * e + 1] = __float2int_rz(f2_hi(p));
            }
            uint32_t o[kW];
#pragma unroll
            for (int w = 0; w < kW; ++w) {
                if (BITS == 4) {
                    uint32_t t = 0;
#pragma unroll
                    for (int b = 3; b >= 0; --b) t = pack_s4(q[8 * w + 2 * b], q[8 * w + 2 * b + 1], t);
                    o[w] = t;
                } else {
                    o[w] = pack_s8(q[4 * w + 1], q[4 * w], pack_s8(q[4 * w + 3], q[4 * w + 2], 0u));
                }
            }
            uint32_t *gr = r + blk * kBW + s * kW;
            if (kW == 2) {
                *reinterpret_cast<uint2 *>(gr) = make_uint2(o[0], o[1]);
            } else {
#pragma unroll
                for (int i = 0; i < kW / 4; ++i) reinterpret_cast<uint4 *>(gr)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < kW; ++i) { wu[i] = nwu[i]; wv[i] = nwv[i]; }
        su_raw = nsu_raw; sv_raw = nsv_raw;
    }
}

template <int BITS, int TPB>
static void launch_scale_add_pk(const uint32_t *u32, const float *su, const uint32_t *v32, const float *sv, float a, uint64_t nblocks,
                                uint32_t *r32, float *sr, cudaStream_t stream) {
    static int ctas_per_sm[kMaxDevices] = {};
    const int dev = current_device() < 0 ? 0 : current_device();
    int &cps = ctas_per_sm[dev];
    if (cps == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, k_vscale_add_pk<BITS, TPB>, 256, 0) != cudaSuccess || cps < 1) cps = 1;
    }
    const uint64_t max_groups = (uint64_t)sm_count() * cps * (256 / TPB);                  // one resident wave
    const uint64_t groups = nblocks < max_groups ? nblocks : max_groups;
    const unsigned grid = (unsigned)((groups * TPB + 255) / 256);
    k_vscale_add_pk<BITS, TPB><<<grid, 256, 0, stream>>>(u32, su, v32, sv, a, nblocks, r32, sr, 0u);
}

template <int BITS>
static int launch_scale_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n_pad,
                            int8_t *r, float *sr, uint64_t *key_host, cudaStream_t stream) {
    const uint64_t nblocks = n_pad / kBlock;
    if (nblocks == 0) return CLOVER_OK;
    const uint32_t *u32 = reinterpret_cast<const uint32_t *>(u), *v32 = reinterpret_cast<const uint32_t *>(v);
    uint32_t *r32 = reinterpret_cast<uint32_t *>(r);
    if (!key_host) {
        // threads per block, measured at n = 2^26 on B200 (r02s): 4-bit 37.5 / 39.7 / 45.7 us with 1 / 2 / 4 threads per block
        // (3.0 TB/s, 46 % of the HBM peak; round 1: 58.7 us), 8-bit 53.5 / 46.2 / 42.9 us (5.0 TB/s, 76 %; round 1: 56 us) -
        // the 8-bit thread-per-block kernel needs 150 registers
        if (BITS == 4) launch_scale_add_pk<BITS, 1>(u32, su, v32, sv, a, nblocks, r32, sr, stream);
        else           launch_scale_add_pk<BITS, 4>(u32, su, v32, sv, a, nblocks, r32, sr, stream);
        count_launch();
        return launch_status("k_vscale_add_pk");
    }
    const uint64_t *tables = device_jump_tables();
    if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
    const Key4 key = key_lanes(key_host);
    static int ctas_per_sm[kMaxDevices] = {};
    const int dev = current_device() < 0 ? 0 : current_device();
    int &cps = ctas_per_sm[dev];
    if (cps == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, k_vscale_add4t<BITS, true>, 256, 0) != cudaSuccess || cps < 1) cps = 1;
    }
    const uint64_t max_groups = (uint64_t)sm_count() * cps * 64;                       // one resident wave
    const uint64_t groups = nblocks < max_groups ? nblocks : max_groups;
    const uint64_t R = (nblocks + groups - 1) / groups;
    const unsigned grid = (unsigned)((groups * 4 + 255) / 256);
    k_vscale_add4t<BITS, true><<<grid, 256, 0, stream>>>(u32, su, v32, sv, a, nblocks, R, r32, sr, key, tables, 0u);
    host_key_skip(key_host, 2 * nblocks);
    count_launch();
    return launch_status("k_vscale_add");
}

}  // namespace clover

using namespace clover;

extern "C" {

int clover_v4_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n_pad,
                            int8_t *r, float *sr, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(u && su && v && sv && r && sr, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(n_pad % 128u == 0, CLOVER_ERR_INVALID, "n_pad must be a multiple of 128 (clover_size_pad)");
    return launch_scale_add<4>(u, su, v, sv, a, n_pad, r, sr, key_host, (cudaStream_t)stream);
}
int clover_v8_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n_pad,
                            int8_t *r, float *sr, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(u && su && v && sv && r && sr, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(n_pad % 128u == 0, CLOVER_ERR_INVALID, "n_pad must be a multiple of 128 (clover_size_pad)");
    return launch_scale_add<8>(u, su, v, sv, a, n_pad, r, sr, key_host, (cudaStream_t)stream);
}

}  // extern "C"
