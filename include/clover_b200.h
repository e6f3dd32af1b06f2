/*
 * clover_b200.h - the C ABI of the B200-native Clover hot path (libclover_b200.so).
 *
 * The reference (astojanov/Clover) has no FFI layer: its boundary is the public method surface of
 * the header-only containers CloverVector4/8 and CloverMatrix4/8. Every entry point below replaces
 * the BODY of one of those methods; the host containers in include/clover_b200/containers.hpp keep
 * the reference's class and method names and forward here.
 *
 * Conventions
 *   - Plain C: pointers and sizes only. All data pointers are DEVICE pointers unless the name says
 *     `_host`. Layouts are byte-identical to the reference's in-memory layouts:
 *       V4: values n_pad/2 bytes (element 2i in the HIGH nibble of byte i), scales n_pad/64 fp32
 *           (include/CloverVector4.h:44-58, 68-103)
 *       V8: values n_pad bytes, scales n_pad/64 fp32                (include/CloverVector8.h:45-78)
 *       M4: row-major nibbles rows*cols/2 bytes, scales[(i>>6)*(cols>>6) + (j>>6)]
 *           (include/CloverMatrix4.h:38-93);  M8: rows*cols bytes, same scales (CloverMatrix8.h:76-92)
 *     n_pad / rows / cols are the PADDED sizes: multiples of 128 (include/CloverVector.h:86-89,
 *     include/CloverMatrix.h:48-50). clover_size_pad() applies the reference's rule.
 *   - Every function returns a clover_status; nothing exits or throws (the reference prints and
 *     exit(1)s, e.g. include/CloverMatrix4.h:779-782 - the host containers keep that behaviour).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream). Calls are
 *     asynchronous with respect to the host unless documented otherwise.
 *   - Stochastic rounding: `key_host` points to the reference's PRNG state, uint64[8] =
 *     random_key1[4] | random_key2[4] (include/CloverRandom.h:39-41), in HOST memory. NULL selects
 *     the reference's CLOVER_STOCHASTIC_ROUNDING_DISABLED behaviour (CMakeLists.txt:78-80). A non-NULL
 *     key is consumed exactly like the reference's sequential code consumes it (same XORShift128+
 *     stream, include/simdxorshift128plus.h:97-109) and is advanced in place before the call returns.
 *   - Streams: calls on DIFFERENT streams of one device are independent - kernel scratch (tickets, fp32 intermediates,
 *     the GEMM's expanded operands) is keyed by (device, stream), allocated on the first call that needs it and grown
 *     with a synchronous cudaMalloc; a replaced block is kept alive, so CUDA graphs captured earlier stay valid. Because
 *     of that first-use allocation, run a call sequence once on the capturing stream BEFORE capturing it. A call with a
 *     non-NULL key cannot be captured usefully (the key is read on the host at launch time).
 *     Calls on the SAME stream are ordered like any other stream work.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CLOVER_ERR_CUDA.
 */
#ifndef CLOVER_B200_H
#define CLOVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum clover_status {
    CLOVER_OK = 0,
    CLOVER_ERR_INVALID = 1,     /* null pointer, size not a multiple of 128, bad flag */
    CLOVER_ERR_CUDA = 2,        /* CUDA runtime error - see clover_last_error() */
    CLOVER_ERR_SIZE = 3,        /* operand sizes do not match (the reference would exit(1)) */
    CLOVER_ERR_UNSUPPORTED = 4
} clover_status;

/* dot() accumulation order. EXACT reproduces the reference SIMD dot's 16 (4-bit)
 * / 8 (8-bit) fp32 FMA chains and its horizontal-add tree bit-for-bit; it is a latency-bound
 * single-warp walk. FAST computes the same exact per-block integers and sums the scaled blocks
 * with a fixed pairwise tree in higher precision (deterministic, not order-identical).
 * AUTO = EXACT up to clover_dot_exact_limit() elements, FAST beyond. */
typedef enum clover_dot_mode { CLOVER_DOT_AUTO = 0, CLOVER_DOT_EXACT = 1, CLOVER_DOT_FAST = 2 } clover_dot_mode;

/* threshold() tie handling. EXACT walks the reference's sequential min-heap (std::make_heap +
 * min_heapify, include/CloverVector4.h:1929-1973) with one thread, so that of several EQUAL magnitudes exactly the
 * reference's survivors remain; FAST is a parallel radix select that keeps all larger magnitudes and, of the elements
 * equal to the k-th largest magnitude, those with the lowest indices. Both leave exactly k non-cleared elements holding
 * the k largest magnitudes (the reference's acceptance test, test/validate/02_vector.cpp:450-500).
 * AUTO = EXACT up to clover_threshold_exact_limit() elements, FAST beyond. */
typedef enum clover_threshold_mode { CLOVER_THRESHOLD_AUTO = 0, CLOVER_THRESHOLD_EXACT = 1, CLOVER_THRESHOLD_FAST = 2 } clover_threshold_mode;

/* ---- library / device --------------------------------------------------------------------------- */
int         clover_version(void);
const char *clover_last_error(void);                 /* thread-local message of the last failure */
int         clover_device_count(void);
int         clover_set_device(int device);
uint64_t    clover_size_pad(uint64_t n);              /* include/CloverVector.h:86-89 */
uint64_t    clover_dot_exact_limit(void);
uint64_t    clover_threshold_exact_limit(void);
int         clover_kernel_launches(void);             /* kernels launched by this library since load */

/* ---- device memory helpers (what the reference does with posix_memalign/free/memcpy) ------------ */
int clover_malloc(void **dev_ptr, size_t bytes);
int clover_free(void *dev_ptr);
int clover_malloc_host(void **host_ptr, size_t bytes);    /* pinned */
int clover_free_host(void *host_ptr);
int clover_memset(void *dev_ptr, int byte, size_t bytes, void *stream);
int clover_copy_h2d(void *dev_dst, const void *host_src, size_t bytes, void *stream);
int clover_copy_d2h(void *host_dst, const void *dev_src, size_t bytes, void *stream);
int clover_copy_d2d(void *dev_dst, const void *dev_src, size_t bytes, void *stream);
int clover_stream_sync(void *stream);

/* ---- peer memory: map another rank's clover_malloc'ed buffer into this process (cudaIpc*), 64-byte handles ----- */
int clover_ipc_export(void *dev_ptr, unsigned char *handle64);
int clover_ipc_import(const unsigned char *handle64, void **dev_ptr);
int clover_ipc_close(void *dev_ptr);

/* ---- PRNG state on the host (include/simdxorshift128plus.h, include/CloverRandom.h) ------------- */
int clover_prng_init(uint64_t key1, uint64_t key2, uint64_t *key_host);   /* avx_xorshift128plus_init :81-92 */
int clover_prng_next(uint64_t *key_host, uint32_t *out8);                 /* avx_xorshift128plus      :97-109 */
int clover_prng_skip(uint64_t *key_host, uint64_t ncalls);                /* O(log n) GF(2) jump-ahead */

/* ---- CloverVector32 / CloverMatrix32: the reference's input generators on the device ------------
 * CloverVector32::setRandomFloats include/CloverVector32.h:751-783, ::setRandomInteger :712-744; the CloverMatrix32 twins
 * (include/CloverMatrix32.h:252-323) run the same loop over size() = rows * cols (PADDED dimensions, the pad is filled
 * too). x: n fp32 in device memory; one XORShift128+ call per 8 elements plus one per left-over element; key_host
 * (required) is consumed exactly like the reference consumes its key pair and advanced in place. These are the
 * synthetic inputs of the reference's harnesses (test/random/00_random.cpp:42, test/performance/01_measure.h:641,711). */
int clover_v32_set_random_floats(float *x, uint64_t n, float min_value, float max_value, uint64_t *key_host, void *stream);
int clover_v32_set_random_integers(float *x, uint64_t n, float min_value, float max_value, uint64_t *key_host, void *stream);

/* ---- CloverVector4 ------------------------------------------------------------------------------ */
/* CloverVector4::quantize      include/CloverVector4.h:605-807   (x: n_pad fp32) */
int clover_v4_quantize(const float *x, uint64_t n_pad, int8_t *values, float *scales, uint64_t *key_host, void *stream);
/* CloverVector4::restore       include/CloverVector4.h:1027-1093 */
int clover_v4_restore(const int8_t *values, const float *scales, uint64_t n_pad, float *x, void *stream);
/* CloverVector4::dot           include/CloverVector4.h:1095-1192 (result: one fp32 in device memory) */
int clover_v4_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n_pad,
                  float *result, int mode, void *stream);

/* CloverVector4::scaleAndAdd   include/CloverVector4.h:1222-1478: r = requantize(u + a * v), block by block.
 * r / sr may alias u / su (the reference's two-argument, in-place form). key_host: the PRNG state of `u`'s object. */
int clover_v4_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n_pad,
                            int8_t *r, float *sr, uint64_t *key_host, void *stream);

/* CloverVector4::threshold(k)  include/CloverVector4.h:1913-1973 (threshold_parallel :1919-1925): keep the k elements
 * of largest magnitude getAbs(i) = |scale/7.0f * q| (:190-203), clear the other nibbles, in place. n = the LOGICAL
 * length (the reference walks `length`, :1931); k >= n leaves the vector unchanged, k = 0 clears it. */
int clover_v4_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k, int mode, void *stream);

/* ---- CloverVector8 ------------------------------------------------------------------------------ */
/* CloverVector8::quantize      include/CloverVector8.h:393-605 */
int clover_v8_quantize(const float *x, uint64_t n_pad, int8_t *values, float *scales, uint64_t *key_host, void *stream);
/* CloverVector8::restore       include/CloverVector8.h:835-909 */
int clover_v8_restore(const int8_t *values, const float *scales, uint64_t n_pad, float *x, void *stream);
/* CloverVector8::dot           include/CloverVector8.h:911-977 */
int clover_v8_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n_pad,
                  float *result, int mode, void *stream);

/* CloverVector8::scaleAndAdd   include/CloverVector8.h:1089-1357 */
int clover_v8_scale_and_add(const int8_t *u, const float *su, const int8_t *v, const float *sv, float a, uint64_t n_pad,
                            int8_t *r, float *sr, uint64_t *key_host, void *stream);

/* CloverVector8::threshold(k)  include/CloverVector8.h:1680-1740; getAbs(i) = |(q * scale) / 127.0f| (:141-147) */
int clover_v8_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k, int mode, void *stream);

/* ---- CloverMatrix4 ------------------------------------------------------------------------------ */
/* CloverMatrix4::quantize      include/CloverMatrix4.h:512-766  (a: rows*cols fp32, row-major) */
int clover_m4_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales,
                       uint64_t *key_host, void *stream);
/* CloverMatrix4::mvm(V4,V4)    include/CloverMatrix4.h:777-1083. y = requantized A*x. `y32` (optional,
 * `rows` fp32) additionally receives the fp32 row results the reference keeps in block_values[]. */
int clover_m4_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                  const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32,
                  uint64_t *key_host, void *stream);
/* CloverMatrix4::mvm(V8,V8)    include/CloverMatrix4.h:1093-1441: mixed precision - the 4-bit matrix times a
 * CloverVector8, result re-quantized to a CloverVector8 (the reference's most accurate IHT configuration). */
int clover_m4_mvm_v8(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                     const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32,
                     uint64_t *key_host, void *stream);
/* CloverMatrix4::mvm(V32,V32)  include/CloverMatrix4.h:1451-1547 */
int clover_m4_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      const float *x32, float *y32, void *stream);
/* CloverMatrix4::restore_scalar include/CloverMatrix4.h:266-301: out (rows * cols fp32) = (scale / 7.0f) * q */
int clover_m4_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out, void *stream);
/* Row-sharded mvm for multi-GPU (no counterpart in the reference): this rank holds rows
 * [row0, row0 + rows_local) of the matrix. Writes the fp32 results into y32_full[row0 ...] of a
 * full-length vector (the NCCL collective runs on that buffer) and the requantized slice into
 * yv/ys at element/row-block offset row0 (so an all-gather of the packed slices is also possible). */
int clover_m4_mvm_shard(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                        uint64_t row0, const int8_t *xv, const float *xs, float *y32_full,
                        int8_t *yv_full, float *ys_full, uint64_t *key_host, void *stream);
/* Fused multi-GPU exchange ("fused NVLink-store epilogue"): like clover_m4_mvm_shard, but the kernel's
 * epilogue stores every finished 64-row block of the re-quantized result (32 B of nibbles + one fp32 scale) straight
 * into the result vector of EVERY rank over NVLink, then the rank's last CTA raises flags[rank] = epoch on all peers and
 * waits for theirs. When the call has completed in stream order, the full CloverVector4 result is present on this
 * rank - no NCCL collective, no separate re-quantize pass. peer_*_host: HOST arrays of `world` device pointers
 * (entry p = rank p's full-length result values / scales and its `world`-word flag array, mapped with
 * clover_ipc_import; entry `rank` = this rank's own buffers). `ticket`: one zero-initialised device word, local.
 * `epoch` must increase by one per call; use two result buffers alternately if a rank may read its result while a
 * faster peer already runs the next call. One node, at most 8 ranks, every rank owns >= 1 row block. */
int clover_m4_mvm_shard_fused(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                              uint64_t row0, const int8_t *xv, const float *xs, int8_t *const *peer_yv_host,
                              float *const *peer_ys_host, uint32_t *const *peer_flags_host, unsigned int *ticket,
                              int world, int rank, uint32_t epoch, uint64_t *key_host, void *stream);
/* Stamped exchange: the same fused epilogue without any ordering between stores. Every 32-bit word of a finished block (8
 * words of nibbles + the scale) travels to every peer in ONE naturally aligned 8-byte store {word, epoch} into that peer's
 * message area (9 x 8 bytes per 64-row block of the WHOLE vector, i.e. rows / 64 * 72 bytes; use two areas alternately,
 * peer_msg_host[p] = rank p's area of this epoch's parity). No system-scope fence per CTA, no flags, the kernel just ends;
 * this rank's own blocks go straight into yv_full / ys_full (reference layout).
 *   clover_m4_shard_stamped_unpack(msg_local, rows, row0, rows_local, epoch, yv_full, ys_full, stream): the consumer side -
 *     polls the stamps of the blocks the OTHER ranks own and writes their words into yv_full / ys_full. When it has completed
 *     in stream order the full CloverVector4 result of call `epoch` is present on this rank.
 * Flow control: peer_started_host[p] = rank p's `world`-word array; a kernel announces "rank r has started call e" on every
 * peer and stores messages only after every peer has started call e too - so whatever consumes the result of call e - 2
 * (same message area) must precede call e in this rank's stream. `epoch` increases by one per call and must be the same on
 * all ranks; zero-initialise message areas and started words once. */
int clover_m4_mvm_shard_stamped(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                                uint64_t row0, const int8_t *xv, const float *xs, int8_t *yv_full, float *ys_full,
                                uint64_t *const *peer_msg_host, uint32_t *const *peer_started_host,
                                int world, int rank, uint32_t epoch, uint64_t *key_host, void *stream);
int clover_m4_shard_stamped_unpack(const uint64_t *msg_local, uint64_t rows, uint64_t row0, uint64_t rows_local, uint32_t epoch,
                                   int8_t *yv_full, float *ys_full, void *stream);
/* Requantize a full-length fp32 vector exactly like the tail of mvm (include/CloverMatrix4.h:925-1080):
 * used after the collective so that every rank holds the same CloverVector4 result. */
int clover_v4_requantize_mvm(const float *y32, uint64_t rows, int8_t *yv, float *ys, uint64_t *key_host, void *stream);
/* 4-bit GEMM (extension): C[i][j] = rowView(A,i).dot(rowView(Bt,j)), A: M x K, Bt: N x K,
 * both CloverMatrix4; C fp32 row-major with leading dimension ldc. */
int clover_m4_gemm(const int8_t *av, const float *as, const int8_t *btv, const float *bts,
                   uint64_t M, uint64_t N, uint64_t K, float *c, uint64_t ldc, void *stream);
/* The two halves of clover_m4_gemm, for callers that reuse an operand (weights): expand the nibbles of a
 * CloverMatrix4 once into rows*cols FP8-E4M3 bytes (integers -8..7 are exact in E4M3) ... */
int clover_m4_expand_e4m3(const int8_t *values, uint64_t rows, uint64_t cols, uint8_t *out, void *stream);
/* ... and run the tcgen05 GEMM on expanded operands (a8: M*K bytes, bt8: N*K bytes, scales as above). */
int clover_m4_gemm_expanded(const uint8_t *a8, const float *as, const uint8_t *bt8, const float *bts,
                            uint64_t M, uint64_t N, uint64_t K, float *c, uint64_t ldc, void *stream);
/* CUDA-core (DP4A) implementation of the same definition; bit-identical to clover_m4_gemm. Validation baseline. */
int clover_m4_gemm_simt(const int8_t *av, const float *as, const int8_t *btv, const float *bts,
                        uint64_t M, uint64_t N, uint64_t K, float *c, uint64_t ldc, void *stream);

/* ---- CloverMatrix8 ------------------------------------------------------------------------------ */
/* CloverMatrix8::quantize      include/CloverMatrix8.h:203-479 */
int clover_m8_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales,
                       uint64_t *key_host, void *stream);
/* CloverMatrix8::mvm(V8,V8)    include/CloverMatrix8.h:1002-1298 */
int clover_m8_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                  const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32,
                  uint64_t *key_host, void *stream);

/* CloverMatrix8::mvm(V32,V32)  include/CloverMatrix8.h:558-661: t = x * (scale / 127.0f), acc = fma(t, q, acc) */
int clover_m8_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      const float *x32, float *y32, void *stream);
/* CloverMatrix8 restore: out[i][j] = get(i, j) = (scale / 127.0f) * q (include/CloverMatrix8.h:117-129; the reference's
 * restore_scalar, :1300-1309, is that assignment inside a loop that does not terminate as written) */
int clover_m8_restore(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, float *out, void *stream);

/* ---- transpose -------------------------------------------------------------------
 * CloverMatrix4::transpose     include/CloverMatrix4.h:1549-1663 (_scalar :435-502, _parallel :2508-2640)
 * CloverMatrix8::transpose     include/CloverMatrix8.h:1359-1385 (_scalar :1312-1336, _parallel :1338-1357)
 * out is the cols x rows matrix with out(j, i) = in(i, j); the 64x64-tile scales are transposed likewise
 * (the reference calls ippiTranspose_32f_C1R for those). Out of place only, like the reference. */
int clover_m4_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                        int8_t *out_values, float *out_scales, void *stream);
int clover_m8_transpose(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                        int8_t *out_values, float *out_scales, void *stream);

/* ---- host-buffer convenience (the call a reference user makes: host containers in, host out) -----
 * Each does pinned-staging H2D, the kernels above, and D2H of the result on `stream`, then syncs. */
int clover_host_v4_quantize(const float *x_host, uint64_t n_pad, int8_t *values_host, float *scales_host, uint64_t *key_host);
int clover_host_v4_dot(const int8_t *u_host, const float *su_host, const int8_t *v_host, const float *sv_host,
                       uint64_t n_pad, float *result_host, int mode);
/* CloverMatrix4::mvm(V4,V4) (include/CloverMatrix4.h:777-1083) for a matrix that is RESIDENT in device memory
 * (uploaded once with clover_copy_h2d, like the reference's matrix object stays in RAM between calls) and per-call
 * vectors in HOST memory: x goes up, the re-quantized y comes back; returns when y is in host memory. Pinned
 * buffers (clover_malloc_host) make the copies asynchronous; a container that is one allocation [values | scales]
 * moves with a single copy per direction. */
int clover_host_m4_mvm(const int8_t *values_dev, const float *scales_dev, uint64_t rows, uint64_t cols,
                       const int8_t *xv_host, const float *xs_host, int8_t *yv_host, float *ys_host, uint64_t *key_host);


/* ---- row-sharded multi-GPU mvm behind one handle (no counterpart in the reference) --------------------
 * ONE host process drives `ngpus` GPUs of the node (devices[] or 0..ngpus-1; peer access must be possible between all of
 * them): rows are sharded in whole 64-row blocks, x is replicated, every GPU runs clover_m4_mvm_shard_fused - the GEMV
 * whose epilogue stores each re-quantized block into every GPU's result vector over NVLink and synchronises with flags,
 * no collective library. The result is the CloverVector4 the single-GPU clover_m4_mvm returns, bit for bit.
 *   _shard      : where rank's rows live (fill them on the device, e.g. with clover_m4_quantize on that device)
 *   _load_host  : scatter a whole matrix (reference layout, host memory) to the shards
 *   _mvm_host   : x from host memory to every GPU, one kernel per GPU, y back to host memory; returns when y is there */
typedef struct clover_m4_sharded clover_m4_sharded;
int clover_m4_sharded_create(clover_m4_sharded **out, uint64_t rows, uint64_t cols, int ngpus, const int *devices);
int clover_m4_sharded_destroy(clover_m4_sharded *h);
int clover_m4_sharded_world(const clover_m4_sharded *h);
int clover_m4_sharded_shard(clover_m4_sharded *h, int rank, int *device, uint64_t *row0, uint64_t *rows_local,
                            int8_t **values_dev, float **scales_dev);
int clover_m4_sharded_load_host(clover_m4_sharded *h, const int8_t *values_host, const float *scales_host);
int clover_m4_sharded_mvm_host(clover_m4_sharded *h, const int8_t *xv_host, const float *xs_host, int8_t *yv_host, float *ys_host,
                               uint64_t *key_host);

#ifdef __cplusplus
}
#endif
#endif /* CLOVER_B200_H */
