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
// One CTA of 256 threads per tile (4-bit: 256x256 elements, 8-bit: 128x128; either way 128 bytes wide on the way in and
// on the way out); 128-bit global loads, register transposes (8x8 nibbles: 16 PRMT + 8 shift/LOP3 pairs; 4x4 bytes: 8 PRMT),
// an XOR-swizzled shared-memory image of the output tile, 128-bit copy-out. See k_mtranspose.
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
// Position-space transpose of W'[k] = in[k ^ 1]; output row c = T[c ^ 1] (both index swaps are free). The 16-bit and
// 8-bit exchange stages are one PRMT per word, the nibble stage a shift and a LOP3 per word: 32 instructions per block.
__device__ __forceinline__ void transpose8x8_nibbles(uint32_t (&a)[8]) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = a[k ^ 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                     // high half of w[i] <-> low half of w[i+4]
        const uint32_t x = w[i], y = w[i + 4];
        w[i] = prmt_b32(x, y, 0x5410u);
        w[i + 4] = prmt_b32(x, y, 0x7632u);
    }
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int i = 4 * g; i < 4 * g + 2; ++i) {     // bytes 1, 3 of w[i] <-> bytes 0, 2 of w[i+2]
            const uint32_t x = w[i], y = w[i + 2];
            w[i] = prmt_b32(x, y, 0x6240u);
            w[i + 2] = prmt_b32(x, y, 0x7351u);
        }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {                  // high nibbles of w[i] <-> low nibbles of w[i+1]
        const uint32_t x = w[i], y = w[i + 1];
        w[i] = (x & 0x0F0F0F0Fu) | ((y << 4) & 0xF0F0F0F0u);
        w[i + 1] = ((x >> 4) & 0x0F0F0F0Fu) | (y & 0xF0F0F0F0u);
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

// A tile is 32 x 32 register blocks of RB x RB elements (RB = 8 nibbles / 4 bytes = one 32-bit word wide): 256 x 256
// elements (4-bit) or 128 x 128 (8-bit), i.e. 128-byte runs on the way in and on the way out. 256 threads.
//   in:   thread (bi, q) reads 16 bytes (word columns 4q..4q+3) of the RB rows of block row bi: RB LDG.128, all in flight;
//   regs: four RB x RB transposes (one per word column);
//   img:  block (bi, bj = 4q + j) becomes words `bi` of output rows R = bj * RB + c. The image keeps 128-byte rows with
//         the 16-byte chunk index XORed by (R / (4 RB)) & 7 = q: the 32 lanes of a warp (4 bi x 8 q) hit 32 banks;
//   out:  thread (row, chunk) copies 16 bytes per iteration: LDS.128 + STG.128, a warp writes four 128-byte runs.
// Round 1 moved single words everywhere (42 issue slots per word, ALU pipe 65 % busy, 71 / 77 % of HBM).
template <int BITS>
__global__ void __launch_bounds__(256)
k_mtranspose(const uint32_t *__restrict__ in, const float *__restrict__ in_scales, uint64_t rows, uint64_t cols,
             uint32_t *__restrict__ out, float *__restrict__ out_scales) {
    constexpr int RB = BITS == 4 ? 8 : 4;             // rows (and columns) per register block
    constexpr int TILE = 32 * RB;                     // 256 / 128
    constexpr int ST = TILE / 64;                     // scale tiles per tile edge
    __shared__ uint4 img4[TILE * 8];
    uint32_t *img = reinterpret_cast<uint32_t *>(img4);

    const int t = threadIdx.x, q = t & 7, bi = t >> 3;
    const uint64_t tiles_i = (rows + TILE - 1) / TILE, tiles_j = (cols + TILE - 1) / TILE, ntiles = tiles_i * tiles_j;
    const uint64_t cpr_in = cols / (4 * RB), cpr_out = rows / (4 * RB);      // 16-byte chunks per row
    const uint64_t vb = rows >> 6, hb = cols >> 6;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    uint4 *out4 = reinterpret_cast<uint4 *>(out);

    // rows and cols are multiples of 128: a 16-byte chunk and a block row are entirely inside or outside
    uint4 v[RB];
    auto load_tile = [&](uint64_t tile) {
        const uint64_t ti = tile / tiles_j, tj = tile % tiles_j;
        const uint64_t row0 = ti * TILE + (uint64_t)bi * RB;
        const bool ok = tj * 8 + q < cpr_in && row0 < rows;
        const uint4 *src = in4 + row0 * cpr_in + tj * 8 + q;
#pragma unroll
        for (int r = 0; r < RB; ++r) v[r] = ok ? ldg_stream(src + (uint64_t)r * cpr_in) : make_uint4(0u, 0u, 0u, 0u);
    };
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t ti = tile / tiles_j, tj = tile % tiles_j;
        load_tile(tile);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t a[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) a[r] = j == 0 ? v[r].x : j == 1 ? v[r].y : j == 2 ? v[r].z : v[r].w;
            if constexpr (BITS == 4) transpose8x8_nibbles(a);
            else                     transpose4x4_bytes(a);
            // output rows R = (4q + j) * RB + c, word bi: chunk (bi >> 2) ^ q, word bi & 3
            uint32_t *dst = img + ((4 * q + j) * RB) * 32 + ((((bi >> 2) ^ q) << 2) | (bi & 3));
#pragma unroll
            for (int c = 0; c < RB; ++c) dst[c * 32] = a[c];
        }
        if (t < ST * ST) {                                        // the tile's 64x64-block scales
            const uint64_t si = ti * ST + t / ST, sj = tj * ST + t % ST;
            if (si < vb && sj < hb) out_scales[sj * vb + si] = in_scales[si * hb + sj];
        }
        __syncthreads();
        const int rr = t >> 3;
        const int64_t rlim = (int64_t)cols - (int64_t)(tj * TILE), clim = (int64_t)cpr_out - (int64_t)(ti * 8);
        uint4 *orow = out4 + (tj * TILE + rr) * cpr_out + ti * 8;
#pragma unroll
        for (int it = 0; it < TILE / 32; ++it) {
            const int R = rr + 32 * it;                           // output row of the tile
            const int ch = q ^ ((R / (4 * RB)) & 7);              // logical chunk held at physical chunk q
            if (ch < clim && R < rlim) __stcs(orow + (uint64_t)it * 32 * cpr_out + ch, img4[R * 8 + q]);
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
    // CTAs per SM the tile loop is spread over, measured on B200 (tools/transpose_sweep.py, 16384^2): 4-bit 47.4 us with 16
    // (49.6 / 53.6 / 50.9 with 3 / 4 / 6), 8-bit 94.5 us with 4 (99.1 / 106 / 106 with 16 / 3 / 6); loading the next tile
    // before the copy-out of the current one (110 registers) was slower for both
    const uint64_t cap = (uint64_t)sm_count() * (BITS == 4 ? 16 : 4);
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
