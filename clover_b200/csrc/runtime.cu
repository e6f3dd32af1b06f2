// runtime.cu - library state: error string, launch counter, PRNG jump tables, memory helpers.
#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include "runtime.cuh"

namespace clover {

static thread_local char g_error[512] = "";
static std::atomic<int> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return CLOVER_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- per-(device, stream, slot) scratch -------------------------------------------------------------
struct ScratchBlock { void *ptr = nullptr; size_t bytes = 0; };
static std::mutex g_scratch_mutex;
static std::map<std::tuple<int, cudaStream_t, int>, ScratchBlock> g_scratch;
static std::vector<void *> g_scratch_retired;         // kept alive until the library is unloaded (captured graphs)

int stream_scratch(ScratchSlot slot, cudaStream_t stream, size_t bytes, size_t zero_bytes, void **out) {
    int dev = 0;
    CLOVER_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    ScratchBlock &b = g_scratch[std::make_tuple(dev, stream, (int)slot)];
    if (b.bytes < bytes) {
        const size_t want = b.bytes ? std::max(bytes, 2 * b.bytes) : bytes;
        void *p = nullptr;
        CLOVER_CUDA_CHECK(cudaMalloc(&p, want));
        if (zero_bytes) {
            // ordered before the caller's first kernel on `stream`; the legacy default stream orders itself
            cudaError_t e = cudaMemsetAsync(p, 0, std::min(zero_bytes, want), stream);
            if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaMemsetAsync(scratch)"); }
        }
        if (b.ptr) g_scratch_retired.push_back(b.ptr);
        b.ptr = p;
        b.bytes = want;
    }
    *out = b.ptr;
    return CLOVER_OK;
}

// ---- jump tables: level i holds L^(2^i) as 8 byte-indexed tables --------------------------------
static std::vector<uint64_t> g_host_tables;
static std::once_flag g_host_tables_once;

static void build_host_tables() {
    g_host_tables.resize((size_t)kJumpLevels * kJumpTableWords);
    uint64_t *t0 = g_host_tables.data();
    for (int p = 0; p < 8; ++p)
        for (int v = 0; v < 256; ++v) t0[p * 256 + v] = xs_advance((uint64_t)v << (8 * p));
    for (int level = 1; level < kJumpLevels; ++level) {
        uint64_t *t = t0 + (size_t)level * kJumpTableWords;
        for (int p = 0; p < 8; ++p)
            for (int v = 0; v < 256; ++v) {
                const uint64_t once = xs_apply_level(t0, level - 1, (uint64_t)v << (8 * p));
                t[p * 256 + v] = xs_apply_level(t0, level - 1, once);
            }
    }
}

const uint64_t *host_jump_tables() {
    std::call_once(g_host_tables_once, build_host_tables);
    return g_host_tables.data();
}

static std::mutex g_dev_tables_mutex;
static uint64_t *g_dev_tables[64] = {nullptr};

const uint64_t *device_jump_tables() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_dev_tables_mutex);
    if (!g_dev_tables[dev]) {
        const uint64_t *h = host_jump_tables();
        const size_t bytes = (size_t)kJumpLevels * kJumpTableWords * sizeof(uint64_t);
        uint64_t *d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
        g_dev_tables[dev] = d;
    }
    return g_dev_tables[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tensor_map_encoder(EncodeTiledFn *out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
        if (q != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled not available in this driver"); return CLOVER_ERR_CUDA; }
        encode = (EncodeTiledFn)fn;
    }
    *out = encode;
    return CLOVER_OK;
}

int make_tensor_map_u8_2d_sw128(CUtensorMap *map, const void *base, uint64_t rows, uint64_t row_bytes,
                                uint32_t box_rows) {
    EncodeTiledFn encode;
    int rc = tensor_map_encoder(&encode);
    if (rc != CLOVER_OK) return rc;
    const cuuint64_t dims[2] = {row_bytes, rows};
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {128, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(sw128) failed with CUresult %d", (int)r); return CLOVER_ERR_CUDA; }
    return CLOVER_OK;
}

int make_tensor_map_u32_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t row_words,
                           uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_words) {
    EncodeTiledFn encode;
    int rc0 = tensor_map_encoder(&encode);
    if (rc0 != CLOVER_OK) return rc0;
    const cuuint64_t dims[2] = {row_words, rows};
    const cuuint64_t strides[1] = {row_pitch_bytes};
    const cuuint32_t box[2] = {box_words, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return CLOVER_ERR_CUDA; }
    return CLOVER_OK;
}

void host_key_skip(uint64_t *key_host, uint64_t ncalls) {
    if (ncalls == 0) return;
    const uint64_t *t = host_jump_tables();
    for (int k = 0; k < 4; ++k) {
        const uint64_t before_last = xs_jump(t, key_host[4 + k], ncalls - 1);
        key_host[k] = before_last;                       // part1 trails part2 by one call
        key_host[4 + k] = xs_advance(before_last);
    }
}

}  // namespace clover

using namespace clover;

extern "C" {

int clover_version(void) { return 100; }
const char *clover_last_error(void) { return g_error; }
int clover_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int clover_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int clover_set_device(int device) { CLOVER_CUDA_CHECK(cudaSetDevice(device)); return CLOVER_OK; }

uint64_t clover_size_pad(uint64_t n) { return (n % 128u) ? n + 128u - (n % 128u) : n; }

int clover_malloc(void **p, size_t bytes) {
    CLOVER_REQUIRE(p != nullptr, CLOVER_ERR_INVALID, "null output pointer");
    CLOVER_CUDA_CHECK(cudaMalloc(p, bytes ? bytes : 1));
    return CLOVER_OK;
}
int clover_free(void *p) { CLOVER_CUDA_CHECK(cudaFree(p)); return CLOVER_OK; }
int clover_malloc_host(void **p, size_t bytes) {
    CLOVER_REQUIRE(p != nullptr, CLOVER_ERR_INVALID, "null output pointer");
    CLOVER_CUDA_CHECK(cudaMallocHost(p, bytes ? bytes : 1));
    return CLOVER_OK;
}
int clover_free_host(void *p) { CLOVER_CUDA_CHECK(cudaFreeHost(p)); return CLOVER_OK; }
int clover_memset(void *p, int byte, size_t bytes, void *stream) {
    CLOVER_CUDA_CHECK(cudaMemsetAsync(p, byte, bytes, (cudaStream_t)stream));
    return CLOVER_OK;
}
int clover_copy_h2d(void *d, const void *h, size_t bytes, void *stream) {
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return CLOVER_OK;
}
int clover_copy_d2h(void *h, const void *d, size_t bytes, void *stream) {
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return CLOVER_OK;
}
int clover_copy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CLOVER_OK;
}
int clover_stream_sync(void *stream) { CLOVER_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream)); return CLOVER_OK; }

// ---- PRNG on the host --------------------------------------------------------------------------
// canonical scalar xorshift128+ step and 2^64 jump: only used to derive lanes 1..3 from lane 0
// (include/simdxorshift128plus.h:38-62, :81-92)
static void scalar_step(uint64_t &s0_slot, uint64_t &s1_slot) {
    uint64_t s1 = s0_slot;
    const uint64_t s0 = s1_slot;
    s0_slot = s0;
    s1 ^= s1 << 23;
    s1_slot = s1 ^ s0 ^ (s1 >> 18) ^ (s0 >> 5);
}
static void scalar_jump(uint64_t in1, uint64_t in2, uint64_t &o1, uint64_t &o2) {
    static const uint64_t poly[2] = {0x8a5cd789635d2dffULL, 0x121fd2155c472f96ULL};
    uint64_t a = 0, b = 0;
    for (int i = 0; i < 2; ++i)
        for (int bit = 0; bit < 64; ++bit) {
            if (poly[i] & (1ULL << bit)) { a ^= in1; b ^= in2; }
            scalar_step(in1, in2);
        }
    o1 = a; o2 = b;
}

int clover_prng_init(uint64_t key1, uint64_t key2, uint64_t *key_host) {
    CLOVER_REQUIRE(key_host != nullptr, CLOVER_ERR_INVALID, "null key");
    key_host[0] = key1; key_host[4] = key2;
    for (int k = 1; k < 4; ++k) scalar_jump(key_host[k - 1], key_host[4 + k - 1], key_host[k], key_host[4 + k]);
    return CLOVER_OK;
}

int clover_prng_next(uint64_t *key_host, uint32_t *out8) {
    CLOVER_REQUIRE(key_host != nullptr && out8 != nullptr, CLOVER_ERR_INVALID, "null pointer");
    for (int k = 0; k < 4; ++k) {
        key_host[k] = key_host[4 + k];
        const uint64_t r = xs_next(key_host[4 + k]);
        out8[2 * k] = (uint32_t)r;
        out8[2 * k + 1] = (uint32_t)(r >> 32);
    }
    return CLOVER_OK;
}

int clover_prng_skip(uint64_t *key_host, uint64_t ncalls) {
    CLOVER_REQUIRE(key_host != nullptr, CLOVER_ERR_INVALID, "null key");
    host_key_skip(key_host, ncalls);
    return CLOVER_OK;
}

// ---- peer memory (CUDA IPC): how one process per GPU maps the other ranks' result vectors for the fused exchange ----
int clover_ipc_export(void *dev_ptr, unsigned char *handle64) {
    CLOVER_REQUIRE(dev_ptr && handle64, CLOVER_ERR_INVALID, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CLOVER_CUDA_CHECK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return CLOVER_OK;
}
int clover_ipc_import(const unsigned char *handle64, void **dev_ptr) {
    CLOVER_REQUIRE(dev_ptr && handle64, CLOVER_ERR_INVALID, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CLOVER_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CLOVER_OK;
}
int clover_ipc_close(void *dev_ptr) {
    CLOVER_REQUIRE(dev_ptr, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return CLOVER_OK;
}

}  // extern "C"
