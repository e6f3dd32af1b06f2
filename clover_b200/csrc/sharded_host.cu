// sharded_host.cu - the row-sharded multi-GPU CloverMatrix4::mvm behind a plain C handle: ONE host process (the
// reference's own model: a C++ program calling container methods) drives all GPUs of the node.
//
// SURVEY.md 8e / 8b (`clover_mvm4_sharded(handles, ...)`): rows are sharded in whole 64-row blocks, x is replicated,
// every GPU runs the GEMV kernel on its rows and its epilogue stores each re-quantized block (8 words of nibbles + one fp32
// scale) into the message area of EVERY GPU through peer pointers (cudaDeviceEnablePeerAccess: NVLink / NVSwitch) as
// self-validating 8-byte {word, epoch} stores - clover_m4_mvm_shard_stamped, one kernel per GPU and call, no NCCL, no flags;
// GPU 0 unpacks the messages into the reference layout in front of the read-back (clover_m4_shard_stamped_unpack). The
// Python host (clover_b200/sharded.py) does the same with one process per GPU and CUDA IPC; the kernels are shared.
#include <vector>
#include "runtime.cuh"

struct clover_m4_sharded {
    uint64_t rows = 0, cols = 0;
    int world = 0;
    uint32_t epoch = 0;
    struct Rank {
        int device = 0;
        cudaStream_t stream = nullptr;
        uint64_t row0 = 0, rows_local = 0;
        int8_t *values = nullptr;        // rows_local * cols / 2
        float *scales = nullptr;         // (rows_local / 64) * (cols / 64)
        unsigned char *x = nullptr;      // [values cols/2 | scales cols/64 fp32]
        unsigned char *block = nullptr;  // [yv x2 | ys x2 | messages x2 | started words], see offsets below
    };
    std::vector<Rank> ranks;
    size_t off_yv[2] = {0, 0}, off_ys[2] = {0, 0}, off_msg[2] = {0, 0}, off_started = 0, block_bytes = 0;
};

namespace {

using namespace clover;

size_t al256(size_t n) { return (n + 255) / 256 * 256; }

struct DeviceGuard {       // the handle's calls leave the caller's current device as they found it
    int saved = 0;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard() { cudaSetDevice(saved); }
};

void shard_rows(uint64_t rows, int world, int rank, uint64_t *row0, uint64_t *rows_local) {   // = sharded.py shard_rows
    const uint64_t blocks = rows / 64, base = blocks / world, extra = blocks % world;
    const uint64_t b0 = (uint64_t)rank * base + ((uint64_t)rank < extra ? (uint64_t)rank : extra);
    *row0 = b0 * 64;
    *rows_local = (base + ((uint64_t)rank < extra ? 1 : 0)) * 64;
}

}  // namespace

extern "C" {

int clover_m4_sharded_create(clover_m4_sharded **out, uint64_t rows, uint64_t cols, int ngpus, const int *devices) {
    CLOVER_REQUIRE(out != nullptr, CLOVER_ERR_INVALID, "null output pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0 && rows > 0 && cols > 0, CLOVER_ERR_INVALID, "rows and cols must be positive multiples of 128");
    CLOVER_REQUIRE(ngpus >= 1 && ngpus <= 8, CLOVER_ERR_INVALID, "1 to 8 GPUs of one node");
    CLOVER_REQUIRE(rows / 64 >= (uint64_t)ngpus, CLOVER_ERR_UNSUPPORTED, "every GPU needs at least one 64-row block");
    int ndev = 0;
    CLOVER_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    DeviceGuard guard;
    clover_m4_sharded *h = new clover_m4_sharded();
    h->rows = rows; h->cols = cols; h->world = ngpus;
    const size_t vb = al256(rows / 2), sb = al256(rows / 64 * sizeof(float));
    h->off_yv[0] = 0; h->off_yv[1] = vb; h->off_ys[0] = 2 * vb; h->off_ys[1] = 2 * vb + sb;
    const size_t mb = al256(rows / 64 * 72);             // 9 x 8 bytes per 64-row block of the whole vector
    h->off_msg[0] = 2 * vb + 2 * sb; h->off_msg[1] = h->off_msg[0] + mb;
    h->off_started = h->off_msg[1] + mb; h->block_bytes = h->off_started + al256(4 * (size_t)ngpus);
    h->ranks.resize(ngpus);
    int rc = CLOVER_OK;
    for (int r = 0; r < ngpus && rc == CLOVER_OK; ++r) {
        clover_m4_sharded::Rank &k = h->ranks[r];
        k.device = devices ? devices[r] : r;
        if (k.device < 0 || k.device >= ndev) { set_error("clover_m4_sharded_create: device %d does not exist", k.device); rc = CLOVER_ERR_INVALID; break; }
        for (int q = 0; q < r; ++q)
            if (h->ranks[q].device == k.device) { set_error("clover_m4_sharded_create: device %d listed twice", k.device); rc = CLOVER_ERR_INVALID; }
        if (rc != CLOVER_OK) break;
        shard_rows(rows, ngpus, r, &k.row0, &k.rows_local);
        auto ok = [&](cudaError_t e, const char *what) { if (e != cudaSuccess && rc == CLOVER_OK) rc = cuda_fail(e, what); return e == cudaSuccess; };
        if (!ok(cudaSetDevice(k.device), "cudaSetDevice")) break;
        ok(cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ok(cudaMalloc(&k.values, k.rows_local * cols / 2), "cudaMalloc(values)");
        ok(cudaMalloc(&k.scales, (k.rows_local / 64) * (cols / 64) * sizeof(float)), "cudaMalloc(scales)");
        ok(cudaMalloc(&k.x, cols / 2 + cols / 64 * sizeof(float)), "cudaMalloc(x)");
        ok(cudaMalloc(&k.block, h->block_bytes), "cudaMalloc(result block)");
        if (rc == CLOVER_OK) ok(cudaMemset(k.block, 0, h->block_bytes), "cudaMemset");
    }
    // every GPU stores into every other GPU's result block
    for (int r = 0; r < ngpus && rc == CLOVER_OK; ++r)
        for (int q = 0; q < ngpus && rc == CLOVER_OK; ++q) {
            if (q == r) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, h->ranks[r].device, h->ranks[q].device);
            if (!can) { set_error("clover_m4_sharded_create: device %d cannot access device %d", h->ranks[r].device, h->ranks[q].device); rc = CLOVER_ERR_UNSUPPORTED; break; }
            cudaSetDevice(h->ranks[r].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(h->ranks[q].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) rc = cuda_fail(e, "cudaDeviceEnablePeerAccess");
        }
    if (rc != CLOVER_OK) { clover_m4_sharded_destroy(h); return rc; }
    *out = h;
    return CLOVER_OK;
}

int clover_m4_sharded_destroy(clover_m4_sharded *h) {
    if (!h) return CLOVER_OK;
    DeviceGuard guard;
    for (auto &k : h->ranks) {
        if (cudaSetDevice(k.device) != cudaSuccess) { cudaGetLastError(); continue; }
        if (k.stream) { cudaStreamSynchronize(k.stream); cudaStreamDestroy(k.stream); }
        cudaFree(k.values); cudaFree(k.scales); cudaFree(k.x); cudaFree(k.block);
    }
    delete h;
    return CLOVER_OK;
}

int clover_m4_sharded_world(const clover_m4_sharded *h) { return h ? h->world : 0; }

int clover_m4_sharded_shard(clover_m4_sharded *h, int rank, int *device, uint64_t *row0, uint64_t *rows_local,
                            int8_t **values_dev, float **scales_dev) {
    CLOVER_REQUIRE(h && rank >= 0 && rank < h->world, CLOVER_ERR_INVALID, "bad handle / rank");
    const clover_m4_sharded::Rank &k = h->ranks[rank];
    if (device) *device = k.device;
    if (row0) *row0 = k.row0;
    if (rows_local) *rows_local = k.rows_local;
    if (values_dev) *values_dev = k.values;
    if (scales_dev) *scales_dev = k.scales;
    return CLOVER_OK;
}

int clover_m4_sharded_load_host(clover_m4_sharded *h, const int8_t *values_host, const float *scales_host) {
    CLOVER_REQUIRE(h && values_host && scales_host, CLOVER_ERR_INVALID, "null pointer");
    DeviceGuard guard;
    const uint64_t hb = h->cols / 64;
    for (auto &k : h->ranks) {
        CLOVER_CUDA_CHECK(cudaSetDevice(k.device));
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(k.values, values_host + k.row0 * h->cols / 2, k.rows_local * h->cols / 2, cudaMemcpyHostToDevice, k.stream));
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(k.scales, scales_host + (k.row0 / 64) * hb, (k.rows_local / 64) * hb * sizeof(float), cudaMemcpyHostToDevice, k.stream));
    }
    for (auto &k : h->ranks) { CLOVER_CUDA_CHECK(cudaSetDevice(k.device)); CLOVER_CUDA_CHECK(cudaStreamSynchronize(k.stream)); }
    return CLOVER_OK;
}

int clover_m4_sharded_mvm_host(clover_m4_sharded *h, const int8_t *xv_host, const float *xs_host, int8_t *yv_host, float *ys_host,
                               uint64_t *key_host) {
    CLOVER_REQUIRE(h && xv_host && xs_host && yv_host && ys_host, CLOVER_ERR_INVALID, "null pointer");
    DeviceGuard guard;
    const uint32_t epoch = ++h->epoch;
    const int b = (int)(epoch & 1);                    // the two result buffers alternate: a fast GPU may already run the next call
    const size_t xvb = h->cols / 2, xsb = h->cols / 64 * sizeof(float);
    uint64_t *peer_msg[8]; uint32_t *peer_started[8];
    for (int r = 0; r < h->world; ++r) {
        unsigned char *blk = h->ranks[r].block;
        peer_msg[r] = reinterpret_cast<uint64_t *>(blk + h->off_msg[b]);
        peer_started[r] = reinterpret_cast<uint32_t *>(blk + h->off_started);
    }
    for (int r = 0; r < h->world; ++r) {
        clover_m4_sharded::Rank &k = h->ranks[r];
        CLOVER_CUDA_CHECK(cudaSetDevice(k.device));
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(k.x, xv_host, xvb, cudaMemcpyHostToDevice, k.stream));
        CLOVER_CUDA_CHECK(cudaMemcpyAsync(k.x + xvb, xs_host, xsb, cudaMemcpyHostToDevice, k.stream));
        // the key is read at each block's GLOBAL position by every GPU and advanced once below, like the reference's stream
        int rc = clover_m4_mvm_shard_stamped(k.values, k.scales, k.rows_local, h->cols, k.row0, reinterpret_cast<const int8_t *>(k.x),
                                             reinterpret_cast<const float *>(k.x + xvb), reinterpret_cast<int8_t *>(k.block + h->off_yv[b]),
                                             reinterpret_cast<float *>(k.block + h->off_ys[b]), peer_msg, peer_started, h->world, r, epoch,
                                             key_host, k.stream);
        if (rc != CLOVER_OK) return rc;
    }
    if (key_host) host_key_skip(key_host, 2 * (h->rows / 64));
    // stamped exchange: no kernel waits for anything at its end; GPU 0 unpacks the other GPUs' messages in front of the read-back
    clover_m4_sharded::Rank &k0 = h->ranks[0];
    CLOVER_CUDA_CHECK(cudaSetDevice(k0.device));
    {
        int rc = clover_m4_shard_stamped_unpack(peer_msg[0], h->rows, k0.row0, k0.rows_local, epoch,
                                                reinterpret_cast<int8_t *>(k0.block + h->off_yv[b]),
                                                reinterpret_cast<float *>(k0.block + h->off_ys[b]), k0.stream);
        if (rc != CLOVER_OK) return rc;
    }
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(yv_host, k0.block + h->off_yv[b], h->rows / 2, cudaMemcpyDeviceToHost, k0.stream));
    CLOVER_CUDA_CHECK(cudaMemcpyAsync(ys_host, k0.block + h->off_ys[b], h->rows / 64 * sizeof(float), cudaMemcpyDeviceToHost, k0.stream));
    for (auto &k : h->ranks) { CLOVER_CUDA_CHECK(cudaSetDevice(k.device)); CLOVER_CUDA_CHECK(cudaStreamSynchronize(k.stream)); }
    return CLOVER_OK;
}

}  // extern "C"
