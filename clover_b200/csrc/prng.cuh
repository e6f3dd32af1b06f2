// prng.cuh - device-side XORShift128+ exactly as the reference's AVX2 code executes it.
//
// Reference: include/simdxorshift128plus.h:97-109. The AVX step overwrites part1 with part2 before
// using it, so each of the four 64-bit lanes is the ONE-word recurrence
//     t = x ^ (x << 23);   x' = t ^ x ^ (t >> 18) ^ (x >> 5);   out = x' + x
// on x = part2 (part1 merely trails one step behind). The map x -> x' is linear over GF(2), which
// gives O(log n) jump-ahead: tables of L^(2^i) as 8 x 256 byte-indexed uint64 entries let any thread
// start at its own position of the sequential stream the reference consumes.
#pragma once
#include <stdint.h>

namespace clover {

constexpr int kJumpLevels = 48;                       // jump distances up to 2^48 calls
constexpr int kJumpTableWords = 8 * 256;              // uint64 entries per level (16 KiB)

__host__ __device__ __forceinline__ uint64_t xs_advance(uint64_t x) {
    const uint64_t t = x ^ (x << 23);
    return t ^ x ^ (t >> 18) ^ (x >> 5);
}

// one call: advances x, returns the 64-bit output (low word = even 32-bit lane, high = odd)
__host__ __device__ __forceinline__ uint64_t xs_next(uint64_t &x) {
    const uint64_t nx = xs_advance(x);
    const uint64_t out = nx + x;
    x = nx;
    return out;
}

// y = L^(2^level) x using the byte tables
__host__ __device__ __forceinline__ uint64_t xs_apply_level(const uint64_t *tables, int level, uint64_t x) {
    const uint64_t *t = tables + (size_t)level * kJumpTableWords;
    uint64_t y = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) y ^= t[p * 256 + ((x >> (8 * p)) & 0xFF)];
    return y;
}

// x <- L^n x
__host__ __device__ __forceinline__ uint64_t xs_jump(const uint64_t *tables, uint64_t x, uint64_t n) {
    for (int level = 0; n != 0 && level < kJumpLevels; ++level, n >>= 1)
        if (n & 1) x = xs_apply_level(tables, level, x);
    return x;
}

}  // namespace clover
