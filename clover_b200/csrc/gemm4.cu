// gemm4.cu - 4-bit GEMM (extension; the reference has no matrix-matrix product).
//
// Definition: C[i][j] = rowView(A,i).dot(rowView(Bt,j)) for A: M x K and Bt: N x K,
// both CloverMatrix4, i.e.  C[i][j] = sum_kb (sA[i>>6][kb] * (1/49) * sB[j>>6][kb]) * I_kb[i][j]
// with an exact int32 partial I_kb per K-slab of 64 and a scale that is constant over each 64x64
// output tile per slab.
//
// k_gemm4_simt: CUDA-core (DP4A) implementation - the parity baseline the tensor-core kernel (gemm4_tc.cu)
// is validated against on the device at sizes the CPU oracle cannot reach. Both accumulate the slabs
// sequentially in fp32 with the same scale expression, so their outputs are bit-identical.
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

__global__ void __launch_bounds__(256)
k_gemm4_simt(const uint32_t *__restrict__ av, const float *__restrict__ as, const uint32_t *__restrict__ bv,
             const float *__restrict__ bs, uint64_t M, uint64_t N, uint64_t K, float *__restrict__ c, uint64_t ldc) {
    // one CTA = one 64x64 tile of C; thread (ty, tx) owns a 4x4 patch
    __shared__ uint32_t a_hi[64][9], a_lo[64][9], b_hi[64][9], b_lo[64][9];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const uint64_t ti = blockIdx.y, tj = blockIdx.x;
    const uint64_t kb_n = K >> 6, wpr = K >> 3;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (uint64_t kb = 0; kb < kb_n; ++kb) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = tid + 256 * h, r = idx >> 3, w = idx & 7;
            const uint32_t wa = av[(ti * 64 + r) * wpr + kb * 8 + w];
            const uint32_t wb = bv[(tj * 64 + r) * wpr + kb * 8 + w];
            a_hi[r][w] = wa & 0xF0F0F0F0u; a_lo[r][w] = (wa << 4) & 0xF0F0F0F0u;
            b_hi[r][w] = wb & 0xF0F0F0F0u; b_lo[r][w] = (wb << 4) & 0xF0F0F0F0u;
        }
        __syncthreads();
        int part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[i][j] = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            int ah[4], al[4], bh[4], bl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ah[i] = (int)a_hi[ty * 4 + i][w]; al[i] = (int)a_lo[ty * 4 + i][w]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { bh[j] = (int)b_hi[tx * 4 + j][w]; bl[j] = (int)b_lo[tx * 4 + j][w]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = dp4a_ss(al[i], bl[j], dp4a_ss(ah[i], bh[j], part[i][j]));
        }
        const float s = __fmul_rn(__fmul_rn(as[ti * kb_n + kb], 1.0f / 49.0f), bs[tj * kb_n + kb]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(s, __int2float_rn(part[i][j] >> 8), acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4 *>(c + (ti * 64 + ty * 4 + i) * ldc + tj * 64 + tx * 4) = o;
    }
}

}  // namespace clover

using namespace clover;

static int gemm_args_ok(const void *av, const void *as, const void *btv, const void *bts, uint64_t M, uint64_t N, uint64_t K,
                        const void *c, uint64_t ldc) {
    CLOVER_REQUIRE(av && as && btv && bts && c, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(M % 128u == 0 && N % 128u == 0 && K % 128u == 0, CLOVER_ERR_INVALID,
                   "M, N, K must be multiples of 128 (CloverMatrix4 padding)");
    CLOVER_REQUIRE(ldc >= N && ldc % 4u == 0, CLOVER_ERR_INVALID, "ldc must be >= N and a multiple of 4");
    CLOVER_REQUIRE(M < (1ull << 31) && N < (1ull << 31) && K < (1ull << 31), CLOVER_ERR_INVALID, "dimension too large");
    return CLOVER_OK;
}

extern "C" {

int clover_m4_gemm(const int8_t *av, const float *as, const int8_t *btv, const float *bts,
                   uint64_t M, uint64_t N, uint64_t K, float *c, uint64_t ldc, void *stream) {
    int rc = gemm_args_ok(av, as, btv, bts, M, N, K, c, ldc);
    if (rc != CLOVER_OK) return rc;
    if (M == 0 || N == 0) return CLOVER_OK;
    return gemm4_tc(av, as, btv, bts, M, N, K, c, ldc, (cudaStream_t)stream);
}

int clover_m4_expand_e4m3(const int8_t *values, uint64_t rows, uint64_t cols, uint8_t *out, void *stream) {
    CLOVER_REQUIRE(values && out, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID, "rows, cols must be multiples of 128");
    return gemm4_expand(values, rows, cols, out, (cudaStream_t)stream);
}

int clover_m4_gemm_expanded(const uint8_t *a8, const float *as, const uint8_t *bt8, const float *bts,
                            uint64_t M, uint64_t N, uint64_t K, float *c, uint64_t ldc, void *stream) {
    int rc = gemm_args_ok(a8, as, bt8, bts, M, N, K, c, ldc);
    if (rc != CLOVER_OK) return rc;
    if (M == 0 || N == 0) return CLOVER_OK;
    return gemm4_tc_expanded(a8, as, bt8, bts, M, N, K, c, ldc, (cudaStream_t)stream);
}

int clover_m4_gemm_simt(const int8_t *av, const float *as, const int8_t *btv, const float *bts,
                   uint64_t M, uint64_t N, uint64_t K, float *c, uint64_t ldc, void *stream) {
    int rc = gemm_args_ok(av, as, btv, bts, M, N, K, c, ldc);
    if (rc != CLOVER_OK) return rc;
    if (M == 0 || N == 0) return CLOVER_OK;
    dim3 grid((unsigned)(N / 64), (unsigned)(M / 64));
    k_gemm4_simt<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(av), as,
                                                         reinterpret_cast<const uint32_t *>(btv), bts, M, N, K, c, ldc);
    count_launch();
    return launch_status("k_gemm4_simt");
}

}  // extern "C"
