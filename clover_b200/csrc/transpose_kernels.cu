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
// One CTA of 256 threads per tile (4-bit: 256x256 elements, 8-bit: 128x128; either way 128 bytes wide on the way in
// and on the way out):
//   1. thread (bi, bj) loads word bj of 8 (4) consecutive rows - a warp reads one 128-byte run per instruction;
//   2. transposes its 8x8 nibbles (three masked-swap stages) / 4x4 bytes (six PRMT) in registers;
//   3. writes the result words into a shared-memory image of the output tile, columns rotated by bj so that the 32
//      lanes of a warp hit 32 different banks;
//   4. the CTA copies the image out row by row: a warp writes one 128-byte run per instruction.
#include <stdlib.h>
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

// A tile is 32 x 32 thread blocks of RB x RB elements (RB = 8 nibbles / 4 bytes = one 32-bit word wide): 256 x 256
// elements (4-bit) or 128 x 128 (8-bit), i.e. 128-byte runs on the way in and on the way out. 256 threads, thread =
// word column bj of four thread-block rows bi: all of its 4 * RB loads are in flight together.
template <int BITS, int RASTER, bool STREAM_ST>
__global__ void __launch_bounds__(256)
k_mtranspose(const uint32_t *__restrict__ in, const float *__restrict__ in_scales, uint64_t rows, uint64_t cols,
             uint32_t *__restrict__ out, float *__restrict__ out_scales) {
    constexpr int RB = BITS == 4 ? 8 : 4;             // rows (and columns) per thread block
    constexpr int TILE = 32 * RB;                     // 256 / 128
    constexpr int ST = TILE / 64;                     // scale tiles per tile edge
    __shared__ uint32_t img[TILE * 32];

    const int t = threadIdx.x, bj = t & 31, bi0 = t >> 5;
    const uint64_t tiles_i = (rows + TILE - 1) / TILE, tiles_j = (cols + TILE - 1) / TILE, ntiles = tiles_i * tiles_j;
    const uint64_t wpr_in = cols / RB, wpr_out = rows / RB;      // words per row (RB elements per word)
    const uint64_t vb = rows >> 6, hb = cols >> 6;

    // RASTER > 1: tiles are walked in RASTER x RASTER super-blocks, so that the CTAs resident at any moment cover square
    // regions: their 128-byte runs line up to RASTER * 128 contiguous bytes per row on the way in AND on the way out
    // (row-major order gives 8 KiB-long reads but isolated 128-byte writes one row pitch apart)
    const uint64_t sb_j = (tiles_j + RASTER - 1) / RASTER, sb_i = (tiles_i + RASTER - 1) / RASTER;
    const uint64_t nwalk = RASTER > 1 ? sb_i * sb_j * RASTER * RASTER : ntiles;
    for (uint64_t walk = blockIdx.x; walk < nwalk; walk += gridDim.x) {
        uint64_t ti, tj;
        if (RASTER > 1) {
            const uint64_t sb = walk / (RASTER * RASTER), in_sb = walk % (RASTER * RASTER);
            ti = (sb / sb_j) * RASTER + in_sb / RASTER;
            tj = (sb % sb_j) * RASTER + in_sb % RASTER;
            if (ti >= tiles_i || tj >= tiles_j) continue;
        } else {
            ti = walk / tiles_j; tj = walk % tiles_j;
        }
        const bool col_ok = tj * 32 + bj < wpr_in;                // rows/cols are multiples of 128: a word is all in or all out
        uint32_t a[4][RB];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t row0 = ti * TILE + (uint64_t)(bi0 + 8 * k) * RB;
            const bool ok = col_ok && row0 < rows;
            const uint32_t *src = in + row0 * wpr_in + tj * 32 + bj;
#pragma unroll
            for (int r = 0; r < RB; ++r) a[k][r] = ok ? ldg_stream(src + (uint64_t)r * wpr_in) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if constexpr (BITS == 4) transpose8x8_nibbles(a[k]);
            else                     transpose4x4_bytes(a[k]);
            const int bi = bi0 + 8 * k;
            // output row (bj * RB + c) of the tile, word bi; columns rotated by bj: the 32 lanes hit 32 banks
#pragma unroll
            for (int c = 0; c < RB; ++c) img[(bj * RB + c) * 32 + ((bi + bj) & 31)] = a[k][c];
        }
        if (t < ST * ST) {                                        // the tile's 64x64-block scales
            const uint64_t si = ti * ST + t / ST, sj = tj * ST + t % ST;
            if (si < vb && sj < hb) out_scales[sj * vb + si] = in_scales[si * hb + sj];
        }
        __syncthreads();
        const bool word_ok = ti * 32 + bj < wpr_out;
        uint32_t *dst = out + (tj * TILE) * wpr_out + ti * 32 + bj;
#pragma unroll 8
        for (int it = 0; it < TILE / 8; ++it) {
            const int r = bi0 + 8 * it;                           // output row of the tile; this thread writes word bj
            if (word_ok && tj * TILE + r < cols) {
                if (STREAM_ST) __stcs(dst + (uint64_t)r * wpr_out, img[r * 32 + ((bj + r / RB) & 31)]);
                else           dst[(uint64_t)r * wpr_out] = img[r * 32 + ((bj + r / RB) & 31)];
            }
        }
        __syncthreads();
    }
}

template <int BITS>
static int launch_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, int8_t *out_values,
                            float *out_scales, cudaStream_t stream) {
    constexpr uint64_t TILE = BITS == 4 ? 256 : 128;
    const uint64_t ntiles = ((rows + TILE - 1) / TILE) * ((cols + TILE - 1) / TILE);
    if (ntiles == 0) return CLOVER_OK;
    const char *v = getenv("CLOVER_TRANSPOSE_IMPL");
    const int variant = v ? atoi(v) : 0;
    const uint64_t R = variant % 10 == 1 ? 4 : variant % 10 == 2 ? 8 : variant % 10 == 3 ? 16 : 1;
    const uint64_t nwalk = R > 1 ? (((rows + TILE - 1) / TILE + R - 1) / R) * (((cols + TILE - 1) / TILE + R - 1) / R) * R * R : ntiles;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    const unsigned grid = (unsigned)(nwalk < cap ? nwalk : cap);
    const uint32_t *in = reinterpret_cast<const uint32_t *>(values);
    uint32_t *out = reinterpret_cast<uint32_t *>(out_values);
    switch (variant) {
        case 1:  k_mtranspose<BITS, 4, true><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        case 2:  k_mtranspose<BITS, 8, true><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        case 3:  k_mtranspose<BITS, 16, true><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        case 10: k_mtranspose<BITS, 1, false><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        case 11: k_mtranspose<BITS, 4, false><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        case 12: k_mtranspose<BITS, 8, false><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        case 13: k_mtranspose<BITS, 16, false><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
        default: k_mtranspose<BITS, 1, true><<<grid, 256, 0, stream>>>(in, scales, rows, cols, out, out_scales); break;
    }
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
