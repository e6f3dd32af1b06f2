// transpose_kernels.cu - CloverMatrix4 / CloverMatrix8 transpose.
//
//   CloverMatrix4::transpose  include/CloverMatrix4.h:1549-1663 (AVX2, 8x8 nibble blocks; scalar :435-502,
//                             parallel :2508-2640, faster scalar :2649-2802)
//   CloverMatrix8::transpose  include/CloverMatrix8.h:1359-1385 (ippiTranspose_8u_C1R; scalar :1312-1336)
//
// A pure permutation: element (i, j) -> (j, i), the 64x64-tile scales are transposed likewise (the reference hands
// those to ippiTranspose_32f_C1R). HBM-bound: every byte is read once and written once (0.5 + 0.5 B/element for
// 4-bit, 1 + 1 for 8-bit, plus the scales).
//
// One CTA of 256 threads per tile (4-bit: 128x128 elements, 8-bit: 64x64; either way 64 bytes wide on the way in
// and on the way out):
//   1. thread (bi, bj) loads word bj of 8 (4) consecutive rows - a warp reads two 64-byte runs per instruction;
//   2. transposes its 8x8 nibbles (three masked-swap stages) / 4x4 bytes (six PRMT) in registers;
//   3. writes the result words into a shared-memory image of the output tile, columns rotated by bj so that the 16
//      lanes of a half-warp hit 16 different banks;
//   4. the CTA copies the image out row by row: a warp writes two 64-byte runs per instruction.
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// a[r] = word of row r holding 8 nibbles in the reference order (element e at nibble position e ^ 1).
// On return a[c] = word of output row c (element r of it = element c of input row r).
__device__ __forceinline__ void transpose8x8_nibbles(uint32_t (&a)[8]) {
    // position-space transpose of W'[k] = in[k ^ 1]; output row c = T[c ^ 1] (both index swaps are free)
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = a[k ^ 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t t = ((w[i] >> 16) ^ w[i + 4]) & 0x0000FFFFu;
        w[i + 4] ^= t; w[i] ^= t << 16;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int i = 4 * g; i < 4 * g + 2; ++i) {
            const uint32_t t = ((w[i] >> 8) ^ w[i + 2]) & 0x00FF00FFu;
            w[i + 2] ^= t; w[i] ^= t << 8;
        }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const uint32_t t = ((w[i] >> 4) ^ w[i + 1]) & 0x0F0F0F0Fu;
        w[i + 1] ^= t; w[i] ^= t << 4;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = w[c ^ 1];
}

__device__ __forceinline__ void transpose4x4_bytes(uint32_t (&a)[4]) {
    const uint32_t t0 = prmt_b32(a[0], a[1], 0x5140u), t1 = prmt_b32(a[0], a[1], 0x7362u);
    const uint32_t u0 = prmt_b32(a[2], a[3], 0x5140u), u1 = prmt_b32(a[2], a[3], 0x7362u);
    a[0] = prmt_b32(t0, u0, 0x5410u);
    a[1] = prmt_b32(t0, u0, 0x7632u);
    a[2] = prmt_b32(t1, u1, 0x5410u);
    a[3] = prmt_b32(t1, u1, 0x7632u);
}

// BITS = 4: tile 128 x 128 elements, thread block 8 x 8 nibbles. BITS = 8: tile 64 x 64, thread block 4 x 4 bytes.
// Either way a tile is 16 x 16 thread blocks, 16 words wide, and RB rows per thread block.
template <int BITS>
__global__ void __launch_bounds__(256)
k_mtranspose(const uint32_t *__restrict__ in, const float *__restrict__ in_scales, uint64_t rows, uint64_t cols,
             uint32_t *__restrict__ out, float *__restrict__ out_scales) {
    constexpr int RB = BITS == 4 ? 8 : 4;             // rows (and columns) per thread block
    constexpr int TILE = 16 * RB;                     // 128 / 64
    constexpr int EPW = BITS == 4 ? 8 : 4;            // elements per 32-bit word
    __shared__ uint32_t img[TILE * 16];

    const int t = threadIdx.x, bi = t >> 4, bj = t & 15;
    const uint64_t tiles_j = cols / TILE, ntiles = (rows / TILE) * tiles_j;
    const uint64_t wpr_in = cols / EPW, wpr_out = rows / EPW;
    const uint64_t vb = rows >> 6, hb = cols >> 6;

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t ti = tile / tiles_j, tj = tile % tiles_j;
        const uint32_t *src = in + (ti * TILE + (uint64_t)bi * RB) * wpr_in + tj * 16 + bj;
        uint32_t a[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) a[r] = ldg_stream(src + (uint64_t)r * wpr_in);
        if constexpr (BITS == 4) transpose8x8_nibbles(a);
        else                     transpose4x4_bytes(a);
        // output row (bj * RB + c) of the tile, word bi; columns rotated by bj
#pragma unroll
        for (int c = 0; c < RB; ++c) img[(bj * RB + c) * 16 + ((bi + bj) & 15)] = a[c];
        // scales of this tile: 2 x 2 entries for the 4-bit tile (128 x 128), one for the 8-bit tile (64 x 64)
        if (BITS == 4) {
            if (t < 4) {
                const uint64_t si = ti * 2 + (t >> 1), sj = tj * 2 + (t & 1);
                out_scales[sj * vb + si] = in_scales[si * hb + sj];
            }
        } else if (t == 0) {
            out_scales[tj * vb + ti] = in_scales[ti * hb + tj];
        }
        __syncthreads();
        uint32_t *dst = out + (tj * TILE) * wpr_out + ti * 16 + bj;
#pragma unroll
        for (int it = 0; it < TILE / 16; ++it) {
            const int r = bi + 16 * it;                                   // output row of the tile; this thread writes word bj
            dst[(uint64_t)r * wpr_out] = img[r * 16 + ((bj + r / RB) & 15)];
        }
        __syncthreads();
    }
}

template <int BITS>
static int launch_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, int8_t *out_values,
                            float *out_scales, cudaStream_t stream) {
    constexpr uint64_t TILE = BITS == 4 ? 128 : 64;
    const uint64_t ntiles = (rows / TILE) * (cols / TILE);
    if (ntiles == 0) return CLOVER_OK;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    k_mtranspose<BITS><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t *>(values), scales, rows, cols,
                                                 reinterpret_cast<uint32_t *>(out_values), out_scales);
    count_launch();
    return launch_status("k_mtranspose");
}

}  // namespace clover

using namespace clover;

extern "C" {

int clover_m4_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, int8_t *out_values,
                        float *out_scales, void *stream) {
    CLOVER_REQUIRE(values && scales && out_values && out_scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID,
                   "rows and cols must be multiples of 128 (include/CloverMatrix.h:48-50)");
    CLOVER_REQUIRE(values != out_values && scales != out_scales, CLOVER_ERR_UNSUPPORTED, "transpose is out of place");
    return launch_transpose<4>(values, scales, rows, cols, out_values, out_scales, (cudaStream_t)stream);
}

int clover_m8_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, int8_t *out_values,
                        float *out_scales, void *stream) {
    CLOVER_REQUIRE(values && scales && out_values && out_scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID,
                   "rows and cols must be multiples of 128 (include/CloverMatrix.h:48-50)");
    CLOVER_REQUIRE(values != out_values && scales != out_scales, CLOVER_ERR_UNSUPPORTED, "transpose is out of place");
    return launch_transpose<8>(values, scales, rows, cols, out_values, out_scales, (cudaStream_t)stream);
}

}  // extern "C"
