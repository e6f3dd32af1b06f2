// generator_kernels.cu - the reference's input generators and the matrix restore, on the device.
//
//   CloverVector32::setRandomFloats / setRandomInteger   include/CloverVector32.h:751-783, :712-744
//   CloverMatrix32::setRandomFloats / setRandomInteger   include/CloverMatrix32.h:289-323, :252-287 (same loop over size())
//   CloverMatrix4::restore_scalar                         include/CloverMatrix4.h:266-301
//   CloverMatrix8 restore = get(i, j) element by element  include/CloverMatrix8.h:117-129 (:1300 loops forever as written)
//
// Generators: one XORShift128+ call (include/simdxorshift128plus.h:97-109) yields eight 32-bit words; element 8c + j of
// the first (n / 8) * 8 elements takes word j of call c, each of the n % 8 left-over elements takes word 0 of a call of
// its own (the reference's masked store, :776-782):  x = fma(float(abs(int32(w))), (max - min) / 2^31, min), the integer
// variant rounds that to the nearest integer (ties to even, _MM_FROUND_TO_NEAREST_INT). abs(INT_MIN) stays INT_MIN
// (_mm256_abs_epi32), so one draw in 2^32 gives min - (max - min), exactly like the reference.
// A thread owns one 64-bit lane of kGenRun consecutive calls: it reaches its position of the sequential stream by one
// GF(2) jump (prng.cuh) and then steps; the host key is advanced by the number of calls consumed.
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

constexpr int kGenRun = 64;          // calls per thread

template <bool INTEGER>
__device__ __forceinline__ float gen_value(uint32_t w, float rcp, float lo) {
    const int32_t v = (int32_t)w;
    const int32_t a = v < 0 ? (int32_t)(0u - (uint32_t)v) : v;           // _mm256_abs_epi32: INT_MIN stays INT_MIN
    const float r = __fmaf_rn(__int2float_rn(a), rcp, lo);
    return INTEGER ? rintf(r) : r;
}

template <bool INTEGER>
__global__ void __launch_bounds__(256)
k_fill_random(float *__restrict__ x, uint64_t n, float rcp, float lo, Key4 key, const uint64_t *__restrict__ tables) {
    const uint64_t nmain = n >> 3, ncalls = nmain + (n & 7);
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = (int)(t & 3);                                           // 64-bit lane = words 2k, 2k+1 of every call
    const uint64_t c0 = (t >> 2) * kGenRun;
    if (c0 >= ncalls) return;
    uint64_t state = xs_jump(tables, key.x[k], c0);
    const uint64_t c1 = c0 + kGenRun < ncalls ? c0 + kGenRun : ncalls;
    for (uint64_t c = c0; c < c1; ++c) {
        const uint64_t o = xs_next(state);
        if (c < nmain) {
            float2 v;
            v.x = gen_value<INTEGER>((uint32_t)o, rcp, lo);
            v.y = gen_value<INTEGER>((uint32_t)(o >> 32), rcp, lo);
            *reinterpret_cast<float2 *>(x + 8 * c + 2 * k) = v;
        } else if (k == 0) {
            x[8 * nmain + (c - nmain)] = gen_value<INTEGER>((uint32_t)o, rcp, lo);
        }
    }
}

static int launch_fill(float *x, uint64_t n, float lo, float hi, uint64_t *key_host, bool integer, cudaStream_t stream) {
    if (n == 0) return CLOVER_OK;
    const uint64_t *tables = device_jump_tables();
    if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
    const float rcp = (hi - lo) / 2147483648.0f;                          // (:759) evaluated in fp32 like the reference
    const uint64_t ncalls = (n >> 3) + (n & 7);
    const uint64_t threads = ((ncalls + kGenRun - 1) / kGenRun) * 4;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    const Key4 key = key_lanes(key_host);
    if (integer) k_fill_random<true><<<grid, 256, 0, stream>>>(x, n, rcp, lo, key, tables);
    else         k_fill_random<false><<<grid, 256, 0, stream>>>(x, n, rcp, lo, key, tables);
    count_launch();
    const int rc = launch_status("k_fill_random");
    if (rc == CLOVER_OK) host_key_skip(key_host, ncalls);
    return rc;
}

// matrix restore: one thread per 32-bit word of values (8 nibbles / 4 bytes)
template <int BITS>
__global__ void __launch_bounds__(256)
k_mrestore(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t rows, uint64_t cols, float *__restrict__ out) {
    constexpr int kPer = BITS == 4 ? 8 : 4;
    constexpr float kQ = BITS == 4 ? 7.0f : 127.0f;
    const uint64_t nwords = rows * cols / kPer, hb = cols >> 6;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = i * kPer, r = e / cols, c = e % cols;
        const float s = __fdiv_rn(__ldg(scales + (r >> 6) * hb + (c >> 6)), kQ);         // scale / 7.0f (:283), / 127.0f (:125)
        const uint32_t w = ldg_stream(values + i);
        float4 *o = reinterpret_cast<float4 *>(out + e);
        if (BITS == 4) {
            float f[8];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int byte = (int)(int8_t)(w >> (8 * b));
                f[2 * b] = __fmul_rn(s, __int2float_rn(byte >> 4));                       // (:293-294)
                f[2 * b + 1] = __fmul_rn(s, __int2float_rn((int)((uint32_t)byte << 28) >> 28));
            }
            o[0] = make_float4(f[0], f[1], f[2], f[3]);
            o[1] = make_float4(f[4], f[5], f[6], f[7]);
        } else {
            o[0] = make_float4(__fmul_rn(s, __int2float_rn((int)(int8_t)w)), __fmul_rn(s, __int2float_rn((int)(int8_t)(w >> 8))),
                               __fmul_rn(s, __int2float_rn((int)(int8_t)(w >> 16))), __fmul_rn(s, __int2float_rn((int)(int8_t)(w >> 24))));
        }
    }
}

template <int BITS>
static int launch_mrestore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out, cudaStream_t stream) {
    const uint64_t nwords = rows * cols / (BITS == 4 ? 8 : 4);
    if (nwords == 0) return CLOVER_OK;
    const uint64_t want = (nwords + 255) / 256, cap = (uint64_t)sm_count() * 16;
    k_mrestore<BITS><<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(reinterpret_cast<const uint32_t *>(values), scales, rows, cols, out);
    count_launch();
    return launch_status("k_mrestore");
}

}  // namespace clover

using namespace clover;

extern "C" {

int clover_v32_set_random_floats(float *x, uint64_t n, float min_value, float max_value, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(x && key_host, CLOVER_ERR_INVALID, "null pointer (the generators always consume a key)");
    return launch_fill(x, n, min_value, max_value, key_host, false, (cudaStream_t)stream);
}
int clover_v32_set_random_integers(float *x, uint64_t n, float min_value, float max_value, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(x && key_host, CLOVER_ERR_INVALID, "null pointer (the generators always consume a key)");
    return launch_fill(x, n, min_value, max_value, key_host, true, (cudaStream_t)stream);
}
int clover_m4_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out, void *stream) {
    CLOVER_REQUIRE(values && scales && out, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID, "rows and cols must be multiples of 128");
    return launch_mrestore<4>(values, scales, rows, cols, out, (cudaStream_t)stream);
}
int clover_m8_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out, void *stream) {
    CLOVER_REQUIRE(values && scales && out, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID, "rows and cols must be multiples of 128");
    return launch_mrestore<8>(values, scales, rows, cols, out, (cudaStream_t)stream);
}

}  // extern "C"
