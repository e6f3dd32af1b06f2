// host_api.cu - host-buffer entry points: the call a user of the reference makes (host containers in,
// host result out). Each stages through grow-only device scratch, launches the same kernels as the
// device-pointer ABI and copies the result back, all on one internal stream, then synchronises.
#include <mutex>
#include "runtime.cuh"

namespace clover {

struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want) {
        if (want <= bytes) return CLOVER_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; bytes = 0;
        CLOVER_CUDA_CHECK(cudaMalloc(&ptr, want));
        bytes = want;
        return CLOVER_OK;
    }
};

static std::mutex g_host_mutex;
static Scratch g_in[2], g_out[2];
static cudaStream_t g_stream = nullptr;

static int host_stream(cudaStream_t *s) {
    if (!g_stream) CLOVER_CUDA_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    *s = g_stream;
    return CLOVER_OK;
}

}  // namespace clover

using namespace clover;

#define CLOVER_TRY(expr) do { int _rc = (expr); if (_rc != CLOVER_OK) return _rc; } while (0)

extern "C" {

int clover_host_v4_quantize(const float *x_host, uint64_t n_pad, int8_t *values_host, float *scales_host,
                            uint64_t *key_host) {
    CLOVER_REQUIRE(x_host && values_host && scales_host, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(n_pad % 128u == 0, CLOVER_ERR_INVALID, "n_pad must be a multiple of 128");
    std::lock_guard<std::mutex> lock(g_host_mutex);
    cudaStream_t s = nullptr;
    CLOVER_TRY(host_stream(&s));
    const size_t xb = n_pad * sizeof(float), vb = n_pad / 2, sb = (n_pad / 64) * sizeof(float);
    CLOVER_TRY(g_in[0].reserve(xb));
    CLOVER_TRY(g_out[0].reserve(vb));
    CLOVER_TRY(g_out[1].reserve(sb));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(g_in[0].ptr, x_host, xb, cudaMemcpyHostToDevice, s));
    CLOVER_TRY(clover_v4_quantize((const float *)g_in[0].ptr, n_pad, (int8_t *)g_out[0].ptr, (float *)g_out[1].ptr, key_host, s));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(values_host, g_out[0].ptr, vb, cudaMemcpyDeviceToHost, s));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(scales_host, g_out[1].ptr, sb, cudaMemcpyDeviceToHost, s));
    CLOVER_CUDA_CHECK(cudaStreamSynchronize(s));
    return CLOVER_OK;
}

int clover_host_v4_dot(const int8_t *u_host, const float *su_host, const int8_t *v_host, const float *sv_host,
                       uint64_t n_pad, float *result_host, int mode) {
    CLOVER_REQUIRE(u_host && su_host && v_host && sv_host && result_host, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(n_pad % 128u == 0, CLOVER_ERR_INVALID, "n_pad must be a multiple of 128");
    std::lock_guard<std::mutex> lock(g_host_mutex);
    cudaStream_t s = nullptr;
    CLOVER_TRY(host_stream(&s));
    const size_t vb = n_pad / 2, sb = (n_pad / 64) * sizeof(float);
    const size_t sb_al = (sb + 255) & ~(size_t)255, vb_al = (vb + 255) & ~(size_t)255;
    CLOVER_TRY(g_in[0].reserve(vb_al + sb_al));
    CLOVER_TRY(g_in[1].reserve(vb_al + sb_al));
    CLOVER_TRY(g_out[0].reserve(256));
    char *du = (char *)g_in[0].ptr, *dv = (char *)g_in[1].ptr;
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(du, u_host, vb, cudaMemcpyHostToDevice, s));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(du + vb_al, su_host, sb, cudaMemcpyHostToDevice, s));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(dv, v_host, vb, cudaMemcpyHostToDevice, s));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(dv + vb_al, sv_host, sb, cudaMemcpyHostToDevice, s));
    CLOVER_TRY(clover_v4_dot((const int8_t *)du, (const float *)(du + vb_al), (const int8_t *)dv,
                             (const float *)(dv + vb_al), n_pad, (float *)g_out[0].ptr, mode, s));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(result_host, g_out[0].ptr, sizeof(float), cudaMemcpyDeviceToHost, s));
    CLOVER_CUDA_CHECK(cudaStreamSynchronize(s));
    return CLOVER_OK;
}

}  // extern "C"
