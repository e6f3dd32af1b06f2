// runtime.cuh - host-side plumbing shared by the ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/clover_b200.h"
#include "prng.cuh"

namespace clover {

struct Key4 { uint64_t x[4]; };                      // the four live 64-bit lanes (part2)

void set_error(const char *fmt, ...);
int  cuda_fail(cudaError_t e, const char *what);     // records the message, returns CLOVER_ERR_CUDA
void count_launch(int n = 1);
int  sm_count();                                      // SMs of the current device (148 on B200)
constexpr int kMaxDevices = 64;
int  current_device();                                // cudaGetDevice, or -1 when outside [0, kMaxDevices)

// Kernel scratch (tickets, partial results, expanded operands): grow-only blocks keyed by (current device, stream, slot),
// so that calls on different streams of one device never share tickets or intermediates (the ABI takes a stream per
// call). A block that has to grow is RETIRED, not freed - a CUDA graph captured earlier may still reference it - and
// the replacement is at least twice as large, which bounds the retired memory by the live block. The first
// `zero_bytes` of a NEW block are zeroed (counters that every kernel leaves clean). Allocation is a synchronous
// cudaMalloc: call once outside a stream capture with the sizes the capture will use (clover_b200.h, "Streams").
enum ScratchSlot { kScratchMvmY = 0, kScratchMvmCounters, kScratchGemm, kScratchThreshold, kScratchDotPartials, kScratchDotTicket, kScratchSlots };
int stream_scratch(ScratchSlot slot, cudaStream_t stream, size_t bytes, size_t zero_bytes, void **out);

// device copy of the jump tables for the CURRENT device (uploaded on first use); nullptr on failure
const uint64_t *device_jump_tables();
const uint64_t *host_jump_tables();

// host key helpers: key_host = part1[4] | part2[4]
inline Key4 key_lanes(const uint64_t *key_host) {
    Key4 k;
    for (int i = 0; i < 4; ++i) k.x[i] = key_host[4 + i];
    return k;
}
void host_key_skip(uint64_t *key_host, uint64_t ncalls);

// Row-major 2D tensor of 32-bit words: `rows` x `row_words`, pitch `row_pitch_bytes` (multiple of 16),
// box = box_rows x box_words, no swizzle, zero fill outside. Encoded through the driver entry point
// (cuTensorMapEncodeTiled), so the library does not link libcuda directly.
int make_tensor_map_u32_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t row_words,
                           uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_words);
// Row-major 2D tensor of bytes, box = box_rows x 128 bytes with the 128-byte shared-memory swizzle that
// tcgen05 K-major operand descriptors (layout type SWIZZLE_128B) expect.
int make_tensor_map_u8_2d_sw128(CUtensorMap *map, const void *base, uint64_t rows, uint64_t row_bytes,
                                uint32_t box_rows);

// gemm4_tc.cu: tensor-core GEMM (expand nibbles to E4M3 in a workspace, then tcgen05) and its two halves
int gemm4_tc(const int8_t *av, const float *as, const int8_t *btv, const float *bts, uint64_t M, uint64_t N, uint64_t K,
             float *c, uint64_t ldc, cudaStream_t stream);
int gemm4_expand(const int8_t *values, uint64_t rows, uint64_t cols, uint8_t *out, cudaStream_t stream);
int gemm4_tc_expanded(const uint8_t *a8, const float *as, const uint8_t *b8, const float *bs, uint64_t M, uint64_t N,
                      uint64_t K, float *c, uint64_t ldc, cudaStream_t stream);

#define CLOVER_CUDA_CHECK(expr)                                             \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return ::clover::cuda_fail(_e, #expr);       \
    } while (0)

#define CLOVER_REQUIRE(cond, code, msg)                                     \
    do {                                                                    \
        if (!(cond)) { ::clover::set_error("%s: %s", __func__, msg); return code; } \
    } while (0)

inline int launch_status(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return CLOVER_OK;
}

}  // namespace clover
