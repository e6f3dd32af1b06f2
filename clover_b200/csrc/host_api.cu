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
// staging scratch and the internal stream are per DEVICE (clover_set_device may switch between calls)
struct HostState { Scratch in[2], out[2]; cudaStream_t stream = nullptr; };
static HostState g_host[64];

static int host_state(HostState **hs) {
    int dev = 0;
    CLOVER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index out of range"); return CLOVER_ERR_INVALID; }
    HostState &h = g_host[dev];
    if (!h.stream) CLOVER_CUDA_CHECK(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    *hs = &h;
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
    HostState *hs = nullptr;
    CLOVER_TRY(host_state(&hs));
    cudaStream_t s = hs->stream;
    Scratch *g_in = hs->in, *g_out = hs->out;
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
    HostState *hs = nullptr;
    CLOVER_TRY(host_state(&hs));
    cudaStream_t s = hs->stream;
    Scratch *g_in = hs->in, *g_out = hs->out;
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


// CloverMatrix4::mvm(V4,V4) with the matrix RESIDENT in device memory (uploaded once - the reference's matrix object
// likewise stays in RAM between calls) and the per-call operands in host memory: x = [values | scales] goes up, the
// re-quantized y comes back, the call returns when y is in host memory.
int clover_host_m4_mvm(const int8_t *values_dev, const float *scales_dev, uint64_t rows, uint64_t cols,
                       const int8_t *xv_host, const float *xs_host, int8_t *yv_host, float *ys_host, uint64_t *key_host) {
    CLOVER_REQUIRE(values_dev && scales_dev && xv_host && xs_host && yv_host && ys_host, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID, "rows and cols must be multiples of 128");
    std::lock_guard<std::mutex> lock(g_host_mutex);
    HostState *hs = nullptr;
    CLOVER_TRY(host_state(&hs));
    cudaStream_t s = hs->stream;
    const size_t xvb = cols / 2, xsb = (cols / 64) * sizeof(float), yvb = rows / 2, ysb = (rows / 64) * sizeof(float);
    const size_t xv_al = (xvb + 255) & ~(size_t)255, yv_al = (yvb + 255) & ~(size_t)255;
    CLOVER_TRY(hs->in[0].reserve(xv_al + xsb));
    CLOVER_TRY(hs->out[0].reserve(yv_al + ysb));
    char *dx = (char *)hs->in[0].ptr, *dy = (char *)hs->out[0].ptr;
    // when the caller's vector is ONE allocation [values | scales] (the reference's layout, include/CloverVector4.h:68-103)
    // and values end on a 256-byte boundary, a single copy per direction moves the whole container
    if ((const char *)xs_host == (const char *)xv_host + xvb && xv_al == xvb) {
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(dx, xv_host, xvb + xsb, cudaMemcpyHostToDevice, s));
    } else {
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(dx, xv_host, xvb, cudaMemcpyHostToDevice, s));
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(dx + xv_al, xs_host, xsb, cudaMemcpyHostToDevice, s));
    }
    // Result buffers in PINNED host memory (clover_malloc_host / cudaHostAlloc: mapped into the device's address space) are
    // written by the kernel's re-quantizer itself - 36 bytes per 64-row block straight over PCIe / C2C - and the device->host
    // copy and its launch latency disappear from every call; pageable buffers take the staged copy below.
    int8_t *yv_map = nullptr;
    float *ys_map = nullptr;
    {
        cudaPointerAttributes av{}, as{};
        if (cudaPointerGetAttributes(&av, yv_host) == cudaSuccess && cudaPointerGetAttributes(&as, ys_host) == cudaSuccess &&
            av.type == cudaMemoryTypeHost && as.type == cudaMemoryTypeHost && av.devicePointer && as.devicePointer) {
            yv_map = static_cast<int8_t *>(av.devicePointer); ys_map = static_cast<float *>(as.devicePointer);
        }
        cudaGetLastError();                                                // a pageable pointer may leave cudaErrorInvalidValue behind
    }
    if (yv_map && ys_map) {
        CLOVER_TRY(clover_m4_mvm(values_dev, scales_dev, rows, cols, (const int8_t *)dx, (const float *)(dx + xv_al),
                                 yv_map, ys_map, nullptr, key_host, s));
        CLOVER_CUDA_CHECK(cudaStreamSynchronize(s));
        return CLOVER_OK;
    }
    CLOVER_TRY(clover_m4_mvm(values_dev, scales_dev, rows, cols, (const int8_t *)dx, (const float *)(dx + xv_al),
                             (int8_t *)dy, (float *)(dy + yv_al), nullptr, key_host, s));
    if ((char *)ys_host == (char *)yv_host + yvb && yv_al == yvb) {
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(yv_host, dy, yvb + ysb, cudaMemcpyDeviceToHost, s));
    } else {
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(yv_host, dy, yvb, cudaMemcpyDeviceToHost, s));
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(ys_host, dy + yv_al, ysb, cudaMemcpyDeviceToHost, s));
    }
    CLOVER_CUDA_CHECK(cudaStreamSynchronize(s));
    return CLOVER_OK;
}

}  // extern "C"
