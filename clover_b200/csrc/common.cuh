// common.cuh - shared device helpers for the clover_b200 kernels (sm_100a).
//
// Everything that decides a bit of the result is written with explicit IEEE intrinsics
// (__fdiv_rn, __fmaf_rn, __fmul_rn, __fadd_rn, __float2int_rz, __int2float_rn): they are never
// contracted, reordered or replaced by approximations, whatever the nvcc flags. The build never
// uses --use_fast_math / -ftz=true.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace clover {

constexpr int kBlock = 64;   // CLOVER_VECTOR_BLOCK, include/CloverVector.h:41

// ---------------------------------------------------------------------------------------------
// quantizer arithmetic (include/CloverVector4.h:644-772, identical in every quantizing routine)
// ---------------------------------------------------------------------------------------------

// zero guard of :661-663 - an absmax whose BIT PATTERN is zero becomes 1.0f
__device__ __forceinline__ float guard_zero(float m) { return (__float_as_uint(m) == 0u) ? 1.0f : m; }

// scale = qmax / max as an IEEE division (:668). Never a reciprocal-multiply: with truncation a
// 1-ulp difference flips the top bucket (15% of blocks quantize their max to 6).
__device__ __forceinline__ float quant_scale(float qmax, float maxv) { return __fdiv_rn(qmax, maxv); }

// q = sign(x) * trunc(fma(|x|, scale, rnd))  (:741-772). trunc is odd-symmetric, so the sign is
// transferred in the float domain (one LOP3) before the single F2I.TRUNC.
__device__ __forceinline__ int quant_one(float x, float scale, float rnd) {
    const float v = __fmaf_rn(fabsf(x), scale, rnd);
    const float s = __uint_as_float(__float_as_uint(v) | (__float_as_uint(x) & 0x80000000u));
    return __float2int_rz(s);
}

// stochastic-rounding noise (:690-734): word w of the PRNG output, byte slot g
__device__ __forceinline__ float noise_from_word(uint32_t w, int g) {
    const uint32_t m = (w & 0x7F7F7F7Fu) << (8 * g);
    return __fmul_rn(__int2float_rn((int)m), 4.656612873077392578125e-10f /* 2^-31 */);
}

// Pack eight quantized values (elements e..e+7, natural order) into one 32-bit word of the
// reference's nibble layout: byte j holds element 2j in its HIGH nibble and 2j+1 in its LOW nibble.
// Horner over the nibble positions (7..0) with plain two's-complement arithmetic, then the
// +8 / ^8 trick turns the borrowed sum back into independent 4-bit two's-complement fields:
//   sum_i q_i * 16^p_i + 0x88888888  has nibble (q_i + 8) at p_i (no borrows since q_i + 8 in [0,15]).
__device__ __forceinline__ uint32_t pack8_nibbles(const int *q) {
    // nibble position p (bit 4p) <- element: p=7:e6, 6:e7, 5:e4, 4:e5, 3:e2, 2:e3, 1:e0, 0:e1
    int w = q[6];
    w = w * 16 + q[7];
    w = w * 16 + q[4];
    w = w * 16 + q[5];
    w = w * 16 + q[2];
    w = w * 16 + q[3];
    w = w * 16 + q[0];
    w = w * 16 + q[1];
    return ((uint32_t)w + 0x88888888u) ^ 0x88888888u;
}

// four int8 values (natural order) into one word
__device__ __forceinline__ uint32_t pack4_bytes(const int *q) {
    return (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) |
           ((uint32_t)q[3] << 24);
}

// ---------------------------------------------------------------------------------------------
// integer dot pieces
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_ss(int a, int b, int c) { return __dp4a(a, b, c); }
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {   // unsigned a, signed b
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// exact sum over the 8 nibble pairs of one word pair: sum h_u*h_v + l_u*l_v
// (the reference's &0xF0 / <<4 trick, include/CloverVector4.h:1134-1181: 16q as int8, products carry 256)
__device__ __forceinline__ int nibble_dot_word(uint32_t u, uint32_t v) {
    const int uh = (int)(u & 0xF0F0F0F0u), ul = (int)((u << 4) & 0xF0F0F0F0u);
    const int vh = (int)(v & 0xF0F0F0F0u), vl = (int)((v << 4) & 0xF0F0F0F0u);
    return dp4a_ss(ul, vl, dp4a_ss(uh, vh, 0)) >> 8;
}

// horizontal-add tree of include/CloverBase.h:149-157 over the 8 lanes held by 8 consecutive
// threads (lane index = low 3 bits): ((a4+a0)+(a6+a2)) + ((a5+a1)+(a7+a3)).
// fp32 addition is commutative bit-for-bit, so a butterfly reproduces the tree on every lane.
__device__ __forceinline__ float hadd8_butterfly(float a, unsigned mask = 0xFFFFFFFFu) {
    a = __fadd_rn(a, __shfl_xor_sync(mask, a, 4));
    a = __fadd_rn(a, __shfl_xor_sync(mask, a, 2));
    a = __fadd_rn(a, __shfl_xor_sync(mask, a, 1));
    return a;
}

__device__ __forceinline__ float warp_max(float m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    return m;
}

// streaming (read-once) global loads: keep them out of L1
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace clover
