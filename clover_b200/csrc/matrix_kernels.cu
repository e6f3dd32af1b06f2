// matrix_kernels.cu - CloverMatrix4 / CloverMatrix8: quantize, mvm (GEMV), mvm with fp32 vectors.
//
//   quantize   include/CloverMatrix4.h:512-766, include/CloverMatrix8.h:203-479
//   mvm(V4,V4) include/CloverMatrix4.h:777-1083      mvm(V8,V8) include/CloverMatrix8.h:1002-1298
//   (mvm with fp32 vectors: mvm_f32_kernels.cu)
//
// The GEMVs are HBM-bound (the matrix is read exactly once: 0.5 B/elem + scales) and reproduce the
// reference's fp32 accumulation ORDER, so the fp32 row results and therefore the re-quantized
// 4/8-bit output vector are bit-identical to the AVX2 code.
#include <stdlib.h>
#include <string.h>
#include "async_copy.cuh"
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

// =============================================================================================
// matrix quantize: one CTA per 64x64 tile, 256 threads, thread (row i, quarter k) owns 16 elements
// =============================================================================================
template <int BITS, bool STOCH>
__global__ void __launch_bounds__(256)
k_mquantize(const float *__restrict__ a, uint64_t rows, uint64_t cols, int8_t *__restrict__ values,
            float *__restrict__ scales, Key4 key, const uint64_t *__restrict__ tables) {
    constexpr float kQmax = BITS == 4 ? 7.0f : 127.0f;
    __shared__ float warp_max_s[8];
    const uint64_t hb = cols >> 6, vb = rows >> 6;
    const uint64_t ntiles = hb * vb;
    const int tid = threadIdx.x, i = tid >> 2, k = tid & 3;

    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t bi = t / hb, bj = t % hb;                 // row-major tile walk (DRAM locality)
        const uint64_t off = ((bi << 6) + i) * cols + (bj << 6) + 16 * k;
        float f[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = ldg_stream(reinterpret_cast<const float4 *>(a + off) + j);
            f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
        }
        float m = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) m = fmaxf(m, fabsf(f[e]));
        m = warp_max(m);
        if ((tid & 31) == 0) warp_max_s[tid >> 5] = m;
        __syncthreads();
        m = warp_max_s[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, warp_max_s[w]);
        __syncthreads();
        m = guard_zero(m);
        if (tid == 0) scales[bi * hb + bj] = m;
        const float scale = quant_scale(kQmax, m);

        // stochastic mode: the reference walks tiles column-block-major (b_j outer, :524-525) and
        // draws two calls per tile row, so row i of this tile sits at call 128*(bj*vb + bi) + 2*i.
        // This thread's 16 elements (16k .. 16k+15 of the row) all belong to call c = k/2.
        uint32_t w8[8];
        if (STOCH) {
            const uint64_t call = 128 * (bj * vb + bi) + 2 * (uint64_t)i + (k >> 1);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint64_t lane = xs_jump(tables, key.x[p], call);
                const uint64_t o = xs_next(lane);
                w8[2 * p] = (uint32_t)o;
                w8[2 * p + 1] = (uint32_t)(o >> 32);
            }
        }
        int q[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            // element 16k+e of the row: byte slot g = ((16k+e)%32)/8 = 2*(k&1) + e/8, word e%8
            const float rnd = STOCH ? noise_from_word(w8[e & 7], 2 * (k & 1) + (e >> 3)) : 0.f;
            q[e] = quant_one(f[e], scale, rnd);
        }
        if (BITS == 4) {
            uint2 o = make_uint2(pack8_nibbles(q), pack8_nibbles(q + 8));
            *reinterpret_cast<uint2 *>(values + (off >> 1)) = o;
        } else {
            uint4 o = make_uint4(pack4_bytes(q), pack4_bytes(q + 4), pack4_bytes(q + 8), pack4_bytes(q + 12));
            *reinterpret_cast<uint4 *>(values + off) = o;
        }
    }
}

template <int BITS>
static int launch_mquantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales,
                            uint64_t *key_host, cudaStream_t stream) {
    const uint64_t ntiles = (rows >> 6) * (cols >> 6);
    if (ntiles == 0) return CLOVER_OK;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    Key4 key = {};
    if (key_host) {
        const uint64_t *tables = device_jump_tables();
        if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
        key = key_lanes(key_host);
        k_mquantize<BITS, true><<<grid, 256, 0, stream>>>(a, rows, cols, values, scales, key, tables);
        host_key_skip(key_host, 128 * ntiles);
    } else {
        k_mquantize<BITS, false><<<grid, 256, 0, stream>>>(a, rows, cols, values, scales, key, nullptr);
    }
    count_launch();
    return launch_status("k_mquantize");
}

// =============================================================================================
// re-quantizer shared by the mvm epilogues (include/CloverMatrix4.h:915-1080, CloverMatrix8.h:1110-1295)
// =============================================================================================
// Called by the first 64 threads of a CTA with y = the fp32 result of row (64*rb + i).
// 4-bit: the reference stores block_values pre-transposed, so element i takes the noise slot
// (call (i%8)/4, byte i%4, word i/8); 8-bit keeps the natural slot (call i/32, byte (i%32)/8, word i%8).
// Multi-GPU fused exchange: the peers' result vectors, mapped into this process (CUDA IPC over
// NVLink). The re-quantizer stores each finished block to every peer as well, the last CTA of the kernel then
// raises flags[peer][rank] = epoch on every peer and waits for the peers' flags - one kernel, no NCCL call.
constexpr int kMaxPeers = 8;
struct PeerOut {
    int world = 1, rank = 0;
    uint32_t epoch = 0;
    unsigned int *ticket = nullptr;          // local: CTAs of this rank that are done
    int8_t *yv[kMaxPeers] = {};              // peer p's result values / scales (entry `rank` unused: yv/ys are local)
    float *ys[kMaxPeers] = {};
    uint32_t *flags[kMaxPeers] = {};         // peer p's flag array, one word per source rank
    int defer = 0;                           // 2 (kXchgStamped): stamped form - no flags, no fence: every word travels with its epoch
    uint64_t *msg[kMaxPeers] = {};           // stamped: peer p's message area of this epoch's parity, 9 x 8 bytes per 64-row block
    uint32_t *started[kMaxPeers] = {};       // stamped: peer p's "rank r has started call e" words (flow control)
};
constexpr int kXchgFlags = 0, kXchgStamped = 2;                   // PeerOut::defer

template <int BITS, bool STOCH>
__device__ __forceinline__ void requantize_block(float y, int i, uint64_t rb, int8_t *__restrict__ yv,
                                                 float *__restrict__ ys, const Key4 &key,
                                                 const uint64_t *__restrict__ tables, float *smem_f, int *smem_q,
                                                 const PeerOut *peers = nullptr) {
    constexpr float kQmax = BITS == 4 ? 7.0f : 127.0f;
    float m = warp_max(fabsf(y));
    if ((i & 31) == 0) smem_f[i >> 5] = m;
    asm volatile("bar.sync 1, 64;");                            // only the 64 epilogue threads
    m = guard_zero(fmaxf(smem_f[0], smem_f[1]));
    if (i == 0) ys[rb] = m;
    const float scale = quant_scale(kQmax, m);
    float rnd = 0.f;
    if (STOCH) {
        int c, g, l;
        if (BITS == 4) { c = (i & 7) >> 2; g = i & 3; l = i >> 3; }
        else           { c = i >> 5; g = (i & 31) >> 3; l = i & 7; }
        uint64_t lane = xs_jump(tables, key.x[l >> 1], 2 * rb + c);
        const uint64_t o = xs_next(lane);
        rnd = noise_from_word((l & 1) ? (uint32_t)(o >> 32) : (uint32_t)o, g);
    }
    smem_q[i] = quant_one(y, scale, rnd);
    asm volatile("bar.sync 1, 64;");
    if (BITS == 4) {
        if (i < 8) {
            const uint32_t w = pack8_nibbles(smem_q + 8 * i);
            reinterpret_cast<uint32_t *>(yv + rb * 32)[i] = w;
            if (peers && peers->defer == kXchgStamped) {
                // stamped exchange: ONE naturally aligned 8-byte store {word, epoch} per peer - data and validity arrive together
                const uint64_t stamped = ((uint64_t)peers->epoch << 32) | w;
                for (int p = 0; p < peers->world; ++p)
                    if (p != peers->rank)
                        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peers->msg[p] + rb * 9 + i), "l"(stamped) : "memory");
            } else if (peers) {
                for (int p = 0; p < peers->world; ++p)
                    if (p != peers->rank) reinterpret_cast<uint32_t *>(peers->yv[p] + rb * 32)[i] = w;     // NVLink store
            }
        }
    } else {
        if (i < 16) reinterpret_cast<uint32_t *>(yv + rb * 64)[i] = pack4_bytes(smem_q + 4 * i);
    }
    if (peers && i == 0) {
        if (peers->defer == kXchgStamped) {
            const uint64_t stamped = ((uint64_t)peers->epoch << 32) | __float_as_uint(m);
            for (int p = 0; p < peers->world; ++p)
                if (p != peers->rank)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peers->msg[p] + rb * 9 + 8), "l"(stamped) : "memory");
        } else {
            for (int p = 0; p < peers->world; ++p)
                if (p != peers->rank) peers->ys[p][rb] = m;
        }
    }
}

// Wait until every peer has raised its flag here to at least `epoch`, then acquire at system scope: the peers' result stores
// before their flags are visible to whatever this warp - and, through barriers, its CTA - does next. The flags live in LOCAL
// memory and are written by the peers over NVLink. Called by ALL 32 lanes of a converged warp: lane q polls peer q's flag, so
// one load instruction covers all peers (one L2 round trip per poll instead of world - 1) and the loop condition is
// warp-uniform. Two things this form avoids on purpose (measured with in-kernel prologue waits in round 2,
// profiles/r02u_exchange_notes.md):
//   * `if (lane == 0) wait(); __syncwarp();` in a warp that works afterwards: the warp left the wait diverged and stayed
//     slow for the rest of the kernel - 273 instead of 201 us per step at 32768 x 65536 on 2 GPUs;
//   * nanosleep back-off: a unit warp that slept ONCE at kernel start slowed the whole kernel by 35 % (182 -> 249 us for a
//     5 us sleep; the same 5 us as a busy-wait on clock64: 184 us).
__device__ __forceinline__ void peer_wait_words(const uint32_t *words, int world, int rank, uint32_t epoch,
                                                unsigned mask = 0xFFFFFFFFu, int lane0 = 0) {     // lanes lane0 .. 31 of the warp take part
    const int q = (int)(threadIdx.x & 31) - lane0;
    const bool mine = q >= 0 && q < world && q != rank;
    const uint32_t *flag = words + (mine ? q : 0);
    bool ok;
    do {
        uint32_t seen = epoch;
        if (mine) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        ok = (int32_t)(seen - epoch) >= 0;
    } while (!__all_sync(mask, ok));
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}
__device__ __forceinline__ void peer_wait_flags(const PeerOut &peers, uint32_t epoch) {
    peer_wait_words(peers.flags[peers.rank], peers.world, peers.rank, epoch);
}

// ---- stamped exchange (peers.defer == kXchgStamped) -------------------------------------------------------------------
// What the flag protocols pay at the end of every kernel is ORDER: a system-scope fence in every CTA between its peer
// stores and its ticket (an NVLink round trip), the last CTA's fence, the flag flight, the poll. The stamped form needs no
// order at all: every 32-bit word of the result (8 words of nibbles + the scale per 64-row block) travels in ONE naturally
// aligned 8-byte store {word, epoch} into a message area on the peer - a single-copy-atomic access, so whoever reads the
// epoch has the word. The producer kernel just ends. The consumer side (k_unpack_stamped, in front of whatever reads the
// result) polls the stamps of the blocks the other ranks own and writes the words into the reference layout.
// Flow control instead of flags: a message area is re-used every second call, so a kernel may only store epoch e once every
// peer has STARTED its call e (whatever consumed epoch e-2 there precedes that call in stream order). The idle lanes
// 1..31 of the TMA issuer warp do both, off every critical path: in CTA 0 they announce the start on all peers, in every CTA
// they poll the local words and then raise a shared-memory flag the epilogue threads check before their first message store
// (~40 us later; a wait in a consumer or unit warp at kernel start costs 2 us per kernel - tools/exchange_probe.py).
__device__ __forceinline__ void stamped_flow_control(const PeerOut &peers, volatile int *started_flag) {   // lanes 1..31 of one warp
    const int p = (int)(threadIdx.x & 31) - 1;
    if (blockIdx.x == 0) {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (p < peers.world && p != peers.rank)
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(peers.started[p] + peers.rank), "r"(peers.epoch) : "memory");
    }
    peer_wait_words(peers.started[peers.rank], peers.world, peers.rank, peers.epoch, 0xFFFFFFFEu, 1);
    if (p == 0) { *started_flag = 1; __threadfence_block(); }
}

// msg: this rank's message area of the epoch's parity; every (block, word) another rank owns is polled until its stamp is
// the epoch, then stored into the reference layout: yv32[rb * 8 + j] (j < 8), ys[rb] (j = 8)
__global__ void __launch_bounds__(256)
k_unpack_stamped(const uint64_t *__restrict__ msg, uint64_t nblocks, uint64_t own0, uint64_t ownn, uint32_t epoch,
                 uint32_t *__restrict__ yv32, float *__restrict__ ys) {
    for (uint64_t idx = (uint64_t)blockIdx.x * 256 + threadIdx.x; idx < nblocks * 9; idx += (uint64_t)gridDim.x * 256) {
        const uint64_t rb = idx / 9, j = idx % 9;
        if (rb >= own0 && rb < own0 + ownn) continue;            // written in place by this rank's own kernel
        uint64_t v;
        do {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(msg + idx) : "memory");
        } while ((uint32_t)(v >> 32) != epoch);
        if (j < 8) yv32[rb * 8 + j] = (uint32_t)v;
        else       ys[rb] = __uint_as_float((uint32_t)v);
    }
}

// Tail of the flag-synchronised fused exchange, one WARP per CTA (all 32 lanes), after the CTA's peer stores have been fenced
// at system scope: take a ticket; the LAST CTA of this rank raises flags[peer][rank] = epoch on every peer (lane p stores to
// peer p) and waits until every peer has raised its flag here - when the kernel ends, the slices of all ranks have landed
// in the local result vector. ONE system-scope fence covers all flag stores (round 1 used st.release.sys per peer: a full
// fence per store, i.e. 7 serialised NVLink round trips at 8 GPUs - the 8 / 19 / 27 us per step of VERDICT r01 weak #4);
// the flags are polled with relaxed loads and acquired once at the end.
__device__ __forceinline__ void peer_signal_and_wait(const PeerOut &peers) {
    const int p = threadIdx.x & 31;
    unsigned int t = 0;
    if (p == 0) t = atomicAdd(peers.ticket, 1u);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    if (t != gridDim.x - 1) return;                              // warp-uniform
    if (p == 0) *peers.ticket = 0u;                              // re-armed for the next launch (stream order)
    __syncwarp();
    asm volatile("fence.acq_rel.sys;" ::: "memory");             // acquires the other CTAs' tickets, releases everything to the peers
    if (p < peers.world && p != peers.rank)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(peers.flags[p] + peers.rank), "r"(peers.epoch) : "memory");
    peer_wait_flags(peers, peers.epoch);
}

// =============================================================================================
// mvm(V4,V4): exact-order 4-bit GEMV
// =============================================================================================
// CTA = one 64-row block (= one block of the output vector), 512 threads.
// thread = (row pair rp, accumulator a, AVX lane l): it IS the reference's fp32 chain (a, l) for rows
// rp and rp+32 and performs one FMA per block pair, in block order (CloverMatrix4.h:813-898).
// Per step it loads the 4 bytes (8 nibbles) of block 2p+a, lane l of each row: 16 threads cover the
// 64 contiguous bytes of a block pair.
//
// Integer part without unpacking to signed nibbles. With h' = h+8, l' = l+8 (one LOP3 each:
// (w ^ 0x88888888) & mask) and the x operand pre-expanded ONCE per CTA and column chunk into
//   xh = plain signed high nibbles, xl16 = 16 * signed low nibbles, cneg = -(786432 + 8*(sum xh + sum xl)),
//   S  = dp4a.u32.s32(16*h', xh) + dp4a.u32.s32(l', xl16)  (accumulated on top of the magic 0x4B400000)
// is 16 * sum(h'*xh + l'*xl) and  fma(as_float(S), 1/16, cneg)  is EXACTLY float(sum h*xh + l*xl):
// no I2F, no shifts, two LOP3 + two DP4A + two FFMA per 8 nibbles.
constexpr int kMvmThreads = 512;
constexpr int kMvmChunkBlocks = 64;          // blocks of x staged per chunk (4096 columns)
constexpr uint32_t kMagicBits = 0x4B400000u; // 12582912.0f = 1.5 * 2^23

struct XUnit { int xh; int xl16; float cneg; float prod; };   // one (block, lane) unit, 16 B

__device__ __forceinline__ int sext_nibbles(uint32_t n) {      // bytes hold 0..15 -> signed -8..7
    return (int)__vsub4(n ^ 0x08080808u, 0x08080808u);
}

template <bool STOCH>
__global__ void __launch_bounds__(kMvmThreads)
k_m4_mvm(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t rows_local, uint64_t cols,
         uint64_t rowblock0, const uint32_t *__restrict__ xv, const float *__restrict__ xs,
         float *__restrict__ y32, int8_t *__restrict__ yv, float *__restrict__ ys, Key4 key,
         const uint64_t *__restrict__ tables) {
    __shared__ __align__(16) XUnit units[kMvmChunkBlocks * 8];
    __shared__ float ysm[64];
    __shared__ float red_f[2];
    __shared__ int red_q[64];

    const int tid = threadIdx.x;
    const int rp = tid >> 4, a = (tid >> 3) & 1, l = tid & 7;
    const uint64_t hb = cols >> 6;
    const uint64_t wpr = cols >> 3;                              // 32-bit words per row

    for (uint64_t rb = blockIdx.x; rb < (rows_local >> 6); rb += gridDim.x) {
        const uint32_t *row0 = values + (rb * 64 + rp) * wpr;
        const uint32_t *row1 = row0 + 32 * wpr;
        const float *su = scales + rb * hb;
        float acc0 = 0.f, acc1 = 0.f;

        for (uint64_t cb = 0; cb < hb; cb += kMvmChunkBlocks) {
            const int nb = (int)((hb - cb) < (uint64_t)kMvmChunkBlocks ? (hb - cb) : kMvmChunkBlocks);
            __syncthreads();                                     // previous chunk fully consumed
            if (tid < nb * 8) {
                const int b = tid >> 3;
                const uint32_t w = xv[(cb + b) * 8 + (tid & 7)];
                const int xh = sext_nibbles((w >> 4) & 0x0F0F0F0Fu);
                const int xl = sext_nibbles(w & 0x0F0F0F0Fu);
                XUnit u;
                u.xh = xh;
                u.xl16 = (int)((w << 4) & 0xF0F0F0F0u);
                u.cneg = -(786432.0f + 8.0f * (float)(dp4a_ss(xh, 0x01010101, 0) + dp4a_ss(xl, 0x01010101, 0)));
                u.prod = __fmul_rn(__fmul_rn(su[cb + b], 1.0f / 49.0f), xs[cb + b]);   // (:834-837)
                units[tid] = u;
            }
            __syncthreads();

            const uint32_t *p0 = row0 + cb * 8 + a * 8 + l;
            const uint32_t *p1 = row1 + cb * 8 + a * 8 + l;
            const int steps = nb >> 1;                           // nb is even (cols % 128 == 0)
            int p = 0;
            for (; p + 4 <= steps; p += 4) {
                uint32_t w0[4], w1[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    w0[s] = ldg_stream(p0 + (p + s) * 16);
                    w1[s] = ldg_stream(p1 + (p + s) * 16);
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const XUnit u = units[((p + s) * 2 + a) * 8 + l];
                    const uint32_t t0 = w0[s] ^ 0x88888888u, t1 = w1[s] ^ 0x88888888u;
                    int s0 = dp4a_us(t0 & 0xF0F0F0F0u, u.xh, (int)kMagicBits);
                    s0 = dp4a_us(t0 & 0x0F0F0F0Fu, u.xl16, s0);
                    int s1 = dp4a_us(t1 & 0xF0F0F0F0u, u.xh, (int)kMagicBits);
                    s1 = dp4a_us(t1 & 0x0F0F0F0Fu, u.xl16, s1);
                    const float i0 = __fmaf_rn(__int_as_float(s0), 0.0625f, u.cneg);
                    const float i1 = __fmaf_rn(__int_as_float(s1), 0.0625f, u.cneg);
                    acc0 = __fmaf_rn(u.prod, i0, acc0);          // (:896-897)
                    acc1 = __fmaf_rn(u.prod, i1, acc1);
                }
            }
            for (; p < steps; ++p) {
                const XUnit u = units[(p * 2 + a) * 8 + l];
                const uint32_t t0 = ldg_stream(p0 + p * 16) ^ 0x88888888u;
                const uint32_t t1 = ldg_stream(p1 + p * 16) ^ 0x88888888u;
                int s0 = dp4a_us(t0 & 0xF0F0F0F0u, u.xh, (int)kMagicBits);
                s0 = dp4a_us(t0 & 0x0F0F0F0Fu, u.xl16, s0);
                int s1 = dp4a_us(t1 & 0xF0F0F0F0u, u.xh, (int)kMagicBits);
                s1 = dp4a_us(t1 & 0x0F0F0F0Fu, u.xl16, s1);
                acc0 = __fmaf_rn(u.prod, __fmaf_rn(__int_as_float(s0), 0.0625f, u.cneg), acc0);
                acc1 = __fmaf_rn(u.prod, __fmaf_rn(__int_as_float(s1), 0.0625f, u.cneg), acc1);
            }
        }
        // acc_1 + acc_2, then the hadd tree (:902-907); the 16 lanes of a row are 16 consecutive threads
        acc0 = __fadd_rn(acc0, __shfl_xor_sync(0xFFFFFFFFu, acc0, 8));
        acc1 = __fadd_rn(acc1, __shfl_xor_sync(0xFFFFFFFFu, acc1, 8));
        acc0 = hadd8_butterfly(acc0);
        acc1 = hadd8_butterfly(acc1);
        if ((tid & 15) == 0) { ysm[rp] = acc0; ysm[rp + 32] = acc1; }
        __syncthreads();
        if (tid < 64) {
            const float y = ysm[tid];
            if (y32) y32[(rowblock0 + rb) * 64 + tid] = y;
            if (yv) requantize_block<4, STOCH>(y, tid, rowblock0 + rb, yv, ys, key, tables, red_f, red_q);
        }
        // ysm / red_* are rewritten only after the next row block's chunk barriers
    }
}


// =============================================================================================
// mvm(V4,V4), pipelined: the same exact-order arithmetic as k_m4_mvm, fed by a TMA ring
// =============================================================================================
// Persistent grid: one CTA per SM walks 64-row blocks (CTA b takes row blocks b, b+grid, ...), so
// 1024 row blocks on 148 SMs cost 7 rounds instead of the 4 half-empty waves of a 2-CTA/SM launch.
//
//   warp 16 (one elected lane)  TMA issuer: per stage two cp.async.bulk.tensor.2d boxes
//                               (32 rows x kKC*32 B of nibbles each) into the ring slot
//   warp 17                     expands the matching slice of x into XUnits (incl. the fp32 scale
//                               product of the block) for the slot, one stage ahead in registers
//   warps 0..15 (512 threads)   consumers: thread = fp32 chain (a, l) of two rows; they wait on the
//                               slot's full-barrier, run kKC/2 FMA steps out of shared memory and
//                               release the slot through its empty-barrier
//
// Memory-level parallelism comes from the ring (kGemvStages x 32 KiB in flight per SM), not from
// registers; the matrix is streamed with an L2 evict-first policy since it is read exactly once.
constexpr int kGemvConsumers = 512;
constexpr int kGemvThreads = kGemvConsumers + 64;
constexpr int kKC = 16;                               // blocks (of 64 columns) per stage
constexpr int kGemvStages = 5;
constexpr int kRowPitch = kKC * 32;                   // dense TMA box rows (512 B)

struct __align__(128) GemvStage {
    uint8_t rows[64 * kRowPitch];                     // box 0: rows 0..31, box 1: rows 32..63
    XUnit units[kKC * 8];
};
struct GemvSmem {
    GemvStage stage[kGemvStages];
    uint64_t full[kGemvStages];
    uint64_t empty[kGemvStages];
    float ysm[64];
    float red_f[2];
    int red_q[64];
    int peers_started;                                // stamped exchange: every peer has started this call
};

__device__ __forceinline__ void gemv4_step(const uint8_t *r0, const uint8_t *r1, const XUnit *units, int blk, int l,
                                           float &acc0, float &acc1) {
    const XUnit u = units[blk * 8 + l];
    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(r0 + blk * 32);
    const uint32_t w1 = *reinterpret_cast<const uint32_t *>(r1 + blk * 32);
    int s0 = dp4a_us(xor_and(w0, 0x88888888u, 0xF0F0F0F0u), u.xh, (int)kMagicBits);
    s0 = dp4a_us(xor_and(w0, 0x88888888u, 0x0F0F0F0Fu), u.xl16, s0);
    int s1 = dp4a_us(xor_and(w1, 0x88888888u, 0xF0F0F0F0u), u.xh, (int)kMagicBits);
    s1 = dp4a_us(xor_and(w1, 0x88888888u, 0x0F0F0F0Fu), u.xl16, s1);
    acc0 = __fmaf_rn(u.prod, __fmaf_rn(__int_as_float(s0), 0.0625f, u.cneg), acc0);     // (:896-897)
    acc1 = __fmaf_rn(u.prod, __fmaf_rn(__int_as_float(s1), 0.0625f, u.cneg), acc1);
}

template <bool STOCH, int XCHG>
__global__ void __launch_bounds__(kGemvThreads, 1)
k_m4_mvm_tma(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ scales, uint64_t rows_local,
             uint64_t cols, uint64_t rowblock0, const uint32_t *__restrict__ xv, const float *__restrict__ xs,
             float *__restrict__ y32, int8_t *__restrict__ yv, float *__restrict__ ys, Key4 key,
             const uint64_t *__restrict__ tables, const __grid_constant__ PeerOut peers) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    GemvSmem &sm = *reinterpret_cast<GemvSmem *>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t hb = cols >> 6, nrb = rows_local >> 6;
    const uint32_t chunks = (uint32_t)((hb + kKC - 1) / kKC);

    if (tid == 0) {
        for (int s = 0; s < kGemvStages; ++s) {
            mbar_init(&sm.full[s], 1 + 32);              // TMA issuer (posts the tx bytes) + the 32 unit lanes
            mbar_init(&sm.empty[s], kGemvConsumers / 32);
        }
        sm.peers_started = 0;
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kGemvConsumers / 32) {
        // ------------------------------- TMA issuer -------------------------------
        if (lane == 0) {
            tma_prefetch_descriptor(&tmap);
            const uint64_t policy = policy_evict_first();
            uint32_t it = 0;
            for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
                for (uint32_t c = 0; c < chunks; ++c, ++it) {
                    const int s = it % kGemvStages;
                    mbar_wait(&sm.empty[s], ((it / kGemvStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&sm.full[s], 64 * kRowPitch);   // boxes are always written in full (OOB = 0)
                    tma_load_2d(sm.stage[s].rows, &tmap, (int)(c * kKC * 8), (int)(rb * 64), &sm.full[s], policy);
                    tma_load_2d(sm.stage[s].rows + 32 * kRowPitch, &tmap, (int)(c * kKC * 8), (int)(rb * 64 + 32),
                                &sm.full[s], policy);
                }
            }
        } else if (XCHG == kXchgStamped && peers.world > 1) {
            stamped_flow_control(peers, &sm.peers_started);   // the issuer warp's idle lanes
        }
    } else if (warp == kGemvConsumers / 32 + 1) {
        // ------------------------------- x-unit warp -------------------------------
        // lane owns units lane, lane+32, lane+64, lane+96 of a stage = (block 4j + lane/8, AVX lane lane%8)
        uint32_t it = 0;
        uint32_t w[4] = {0, 0, 0, 0};
        float sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
        auto prefetch = [&](uint64_t rb, uint32_t c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t b = (uint64_t)c * kKC + 4 * j + (lane >> 3);
                const bool ok = b < hb;
                w[j] = ok ? __ldg(xv + b * 8 + (lane & 7)) : 0u;
                sa[j] = ok ? __ldg(scales + rb * hb + b) : 0.f;
                sb[j] = ok ? __ldg(xs + b) : 0.f;
            }
        };
        if (blockIdx.x < nrb) prefetch(blockIdx.x, 0);
        for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            for (uint32_t c = 0; c < chunks; ++c, ++it) {
                const int s = it % kGemvStages;
                XUnit u[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xh = sext_nibbles((w[j] >> 4) & 0x0F0F0F0Fu);
                    const int xl = sext_nibbles(w[j] & 0x0F0F0F0Fu);
                    u[j].xh = xh;
                    u[j].xl16 = (int)((w[j] << 4) & 0xF0F0F0F0u);
                    u[j].cneg = -(786432.0f + 8.0f * (float)(dp4a_ss(xh, 0x01010101, 0) + dp4a_ss(xl, 0x01010101, 0)));
                    u[j].prod = __fmul_rn(__fmul_rn(sa[j], 1.0f / 49.0f), sb[j]);              // (:834-837)
                }
                // next stage's operands are requested before we block on the slot
                uint32_t nc = c + 1; uint64_t nrbi = rb;
                if (nc == chunks) { nc = 0; nrbi = rb + gridDim.x; }
                if (nrbi < nrb) prefetch(nrbi, nc);
                mbar_wait(&sm.empty[s], ((it / kGemvStages) & 1) ^ 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) sm.stage[s].units[lane + 32 * j] = u[j];
                mbar_arrive(&sm.full[s]);
            }
        }
    } else {
        // ------------------------------- consumer warps ------------------------------
        // warp w, half-warp h: rows (32h + w) and (32h + w + 16) of the block; a = accumulator, l = AVX lane
        const int h = lane >> 4, a = (lane >> 3) & 1, l = lane & 7;
        const int row_a = 32 * h + warp, row_b = row_a + 16;
        uint32_t it = 0;
        for (uint64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
            float acc0 = 0.f, acc1 = 0.f;
            for (uint32_t c = 0; c < chunks; ++c, ++it) {
                const int s = it % kGemvStages;
                mbar_wait(&sm.full[s], (it / kGemvStages) & 1);
                const GemvStage &st = sm.stage[s];
                const uint8_t *r0 = st.rows + row_a * kRowPitch + 4 * l;
                const uint8_t *r1 = st.rows + row_b * kRowPitch + 4 * l;
                const uint64_t cb = (uint64_t)c * kKC;
                const int nb = (int)((hb - cb) < (uint64_t)kKC ? (hb - cb) : kKC);
                if (nb == kKC) {
#pragma unroll
                    for (int p = 0; p < kKC / 2; ++p) gemv4_step(r0, r1, st.units, 2 * p + a, l, acc0, acc1);
                } else {
                    for (int p = 0; p < nb / 2; ++p) gemv4_step(r0, r1, st.units, 2 * p + a, l, acc0, acc1);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
            }
            acc0 = __fadd_rn(acc0, __shfl_xor_sync(0xFFFFFFFFu, acc0, 8));       // acc_1 + acc_2, hadd tree (:902-907)
            acc1 = __fadd_rn(acc1, __shfl_xor_sync(0xFFFFFFFFu, acc1, 8));
            acc0 = hadd8_butterfly(acc0);
            acc1 = hadd8_butterfly(acc1);
            named_bar_sync(2, kGemvConsumers);                                   // previous epilogue done with ysm
            if ((lane & 15) == 0) { sm.ysm[row_a] = acc0; sm.ysm[row_b] = acc1; }
            named_bar_sync(2, kGemvConsumers);
            if (tid < 64) {
                const float y = sm.ysm[tid];
                if (y32) y32[(rowblock0 + rb) * 64 + tid] = y;
                if (XCHG == kXchgStamped && peers.world > 1) { while (*(volatile int *)&sm.peers_started == 0) { } }
                if (yv) requantize_block<4, STOCH>(y, tid, rowblock0 + rb, yv, ys, key, tables, sm.red_f, sm.red_q,
                                                   peers.world > 1 ? &peers : nullptr);
            }
        }
        if (XCHG != kXchgStamped && peers.world > 1) {
            if (tid < 64) __threadfence_system();                // this CTA's peer stores are visible system-wide ...
            named_bar_sync(2, kGemvConsumers);                   // ... before its ticket is taken
            if (warp == 0) peer_signal_and_wait(peers);        // all 32 lanes of consumer warp 0
        }
    }
}

// =============================================================================================
// mvm(V8,V8): exact-order GEMV with an 8-bit product vector. 8 chains per row (one accumulator,
// CloverMatrix8.h:1029-1095); thread = (row, lane l): per block the int32 lane sum covers elements 4l..4l+3 and
// 32+4l..32+4l+3. MBITS = 8: CloverMatrix8::mvm (CloverMatrix8.h:1002-1298). MBITS = 4: the mixed-precision
// CloverMatrix4::mvm(V8,V8) (CloverMatrix4.h:1093-1441) - same chains on a nibble matrix, scale
// (su * (1/7)) * (sv * (1/127)), nibbles expanded to 16*q bytes exactly like the reference and the factor 16
// removed exactly in the int -> float step.
// =============================================================================================
constexpr int kMvm8ChunkBlocks = 64;

// two bytes (four nibbles: elements e, e+1 in byte 0, e+2, e+3 in byte 1) -> [16*q_e, 16*q_e+1, 16*q_e+2, 16*q_e+3]
// (three instructions: h << 4 puts every LOW nibble into the high half of a byte, one PRMT interleaves the bytes of h
// and of h << 4, one AND drops the neighbours' bits)
__device__ __forceinline__ uint32_t nibbles4_to_bytes16(uint32_t h) {
    uint32_t p;
    asm("prmt.b32 %0, %1, %2, 0x5140;" : "=r"(p) : "r"(h), "r"(h << 4));
    return p & 0xF0F0F0F0u;
}

template <int MBITS, bool STOCH>
__global__ void __launch_bounds__(512)
k_m8_mvm(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t rows_local, uint64_t cols,
         uint64_t rowblock0, const uint32_t *__restrict__ xv, const float *__restrict__ xs,
         float *__restrict__ y32, int8_t *__restrict__ yv, float *__restrict__ ys, Key4 key,
         const uint64_t *__restrict__ tables) {
    __shared__ uint32_t xsm[kMvm8ChunkBlocks * 16];
    __shared__ float prod[kMvm8ChunkBlocks];
    __shared__ float ysm[64];
    __shared__ float red_f[2];
    __shared__ int red_q[64];

    const int tid = threadIdx.x;
    const int r = tid >> 3, l = tid & 7;
    const uint64_t hb = cols >> 6;
    const uint64_t wpr = MBITS == 8 ? cols >> 2 : cols >> 3;         // 32-bit words per matrix row
    constexpr int kBlockWords = MBITS == 8 ? 16 : 8;

    for (uint64_t rb = blockIdx.x; rb < (rows_local >> 6); rb += gridDim.x) {
        const uint32_t *row = values + (rb * 64 + r) * wpr;
        const float *su = scales + rb * hb;
        float acc = 0.f;
        for (uint64_t cb = 0; cb < hb; cb += kMvm8ChunkBlocks) {
            const int nb = (int)((hb - cb) < (uint64_t)kMvm8ChunkBlocks ? (hb - cb) : kMvm8ChunkBlocks);
            __syncthreads();
            for (int i = tid; i < nb * 16; i += 512) xsm[i] = xv[cb * 16 + i];
            if (tid < nb) {
                if (MBITS == 8) prod[tid] = __fmul_rn(__fmul_rn(su[cb + tid], 1.0f / 127.0f), __fmul_rn(xs[cb + tid], 1.0f / 127.0f));
                else            prod[tid] = __fmul_rn(__fmul_rn(su[cb + tid], 1.0f / 7.0f), __fmul_rn(xs[cb + tid], 1.0f / 127.0f));
            }
            __syncthreads();
            if (MBITS == 8) {
                const uint32_t *p = row + cb * 16 + l;
                int b = 0;
                for (; b + 4 <= nb; b += 4) {
                    uint32_t w0[4], w1[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        w0[s] = ldg_stream(p + (b + s) * 16);
                        w1[s] = ldg_stream(p + (b + s) * 16 + 8);
                    }
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        int d = dp4a_ss((int)w0[s], (int)xsm[(b + s) * 16 + l], 0);
                        d = dp4a_ss((int)w1[s], (int)xsm[(b + s) * 16 + 8 + l], d);
                        acc = __fmaf_rn(prod[b + s], __int2float_rn(d), acc);
                    }
                }
                for (; b < nb; ++b) {
                    int d = dp4a_ss((int)ldg_stream(p + b * 16), (int)xsm[b * 16 + l], 0);
                    d = dp4a_ss((int)ldg_stream(p + b * 16 + 8), (int)xsm[b * 16 + 8 + l], d);
                    acc = __fmaf_rn(prod[b], __int2float_rn(d), acc);
                }
            } else {
                // block = 32 bytes of nibbles; lane l owns bytes 2l, 2l+1 (elements 4l..4l+3) and 16+2l, 17+2l
                const uint16_t *p = reinterpret_cast<const uint16_t *>(row + cb * kBlockWords) + l;
                int b = 0;
                for (; b + 4 <= nb; b += 4) {
                    uint32_t h0[4], h1[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        h0[s] = __ldg(p + (b + s) * 16);
                        h1[s] = __ldg(p + (b + s) * 16 + 8);
                    }
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        int d = dp4a_ss((int)nibbles4_to_bytes16(h0[s]), (int)xsm[(b + s) * 16 + l], 0);
                        d = dp4a_ss((int)nibbles4_to_bytes16(h1[s]), (int)xsm[(b + s) * 16 + 8 + l], d);
                        acc = __fmaf_rn(prod[b + s], __int2float_rn(d >> 4), acc);      // srai 4 (:1196), exact
                    }
                }
                for (; b < nb; ++b) {
                    int d = dp4a_ss((int)nibbles4_to_bytes16(__ldg(p + b * 16)), (int)xsm[b * 16 + l], 0);
                    d = dp4a_ss((int)nibbles4_to_bytes16(__ldg(p + b * 16 + 8)), (int)xsm[b * 16 + 8 + l], d);
                    acc = __fmaf_rn(prod[b], __int2float_rn(d >> 4), acc);
                }
            }
        }
        acc = hadd8_butterfly(acc);
        if (l == 0) ysm[r] = acc;
        __syncthreads();
        if (tid < 64) {
            const float y = ysm[tid];
            if (y32) y32[(rowblock0 + rb) * 64 + tid] = y;
            if (yv) requantize_block<8, STOCH>(y, tid, rowblock0 + rb, yv, ys, key, tables, red_f, red_q);
        }
    }
}

// =============================================================================================
// mvm(V8,V8), pipelined (BASELINE C5): the arithmetic of k_m8_mvm<8> fed by a TMA ring, with 32-row work items
// =============================================================================================
// Why a second kernel: k_m8_mvm launches one CTA per 64-row block and streams with plain loads - 512 row blocks on
// 148 SMs run as 3.46 "rounds" (86 % ceiling) and the loads of a thread are 64 bytes apart. Here
//   * the grid is persistent (one CTA per SM) and a work item is HALF a row block (32 rows): 1024 items at
//     32768 rows = 6.92 rounds (98.8 %). The two halves of a block may be computed by different CTAs: each writes
//     its 32 fp32 row results to global memory, takes a ticket on the block's counter, and the second finisher
//     re-quantizes the 64 values (threadfence + atomic "last block" pattern) and re-arms the counter;
//   * warp 8 (one lane) issues, per stage, eight cp.async.bulk.tensor.2d boxes of 32 rows x 128 B with the
//     128-byte shared-memory swizzle into a 5-stage ring (160 KiB in flight per SM), L2 evict-first;
//   * warp 9 expands the matching 16 blocks of x into (xa, xb, prod) units;
//   * warps 0-7: thread = the reference's fp32 chain (row r, AVX lane l) (CloverMatrix8.h:1029-1095). Warp =
//     8 rows x 4 lanes (l = 4*type .. 4*type+3): the swizzle XORs the 16-byte chunk index with (row & 7), so the
//     32 words of a warp-wide load fall into 32 different banks. Per block: 2 LDS.32 + 1 LDS.128 + 2 DP4A on top
//     of the magic constant 1.5*2^23 (as_float(sum) - 1.5*2^23 IS float(sum): no I2F) + 1 FADD + 1 FFMA, FMAs in
//     block order => fp32 row results and re-quantized bytes identical to k_m8_mvm and to the AVX2 code.
constexpr int kG8Rows = 32;                 // rows per work item
constexpr int kG8Chunks = 8;                // 128-byte column chunks per stage = 16 blocks of 64 columns
constexpr int kG8Consumers = 256;
constexpr int kG8Threads = kG8Consumers + 64;
constexpr int kG8ChunkBytes = kG8Rows * 128;

// a 128-byte chunk holds 2 blocks of 64 columns
struct __align__(1024) Gemv8Stage {
    static constexpr int kBlocksPerChunk = 2;
    uint8_t rows[kG8Chunks][kG8ChunkBytes];   // chunk c: rows 0..31 x 128 B, SWIZZLE_128B (each 4 KiB, 1024-aligned)
    uint4 units[kG8Chunks * kBlocksPerChunk * 8];   // unit (block, l): x = xa, y = xb, z = bits of prod
};
// STAGES = 5: one CTA per SM; STAGES = 3 (105 KiB): two CTAs per SM, grid = 2 x SMs (see Gemv4Smem)
template <int STAGES>
struct Gemv8Smem {
    Gemv8Stage stage[STAGES];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    float part[kG8Rows][8];
    float ysm[64];
    float red_f[2];
    int red_q[64];
    unsigned int ticket;
};

template <bool STOCH, int STAGES>
__global__ void __launch_bounds__(kG8Threads, STAGES == 5 ? 1 : 2)
k_m8_mvm_tma(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ scales, uint64_t rows_local,
             uint64_t cols, uint64_t rowblock0, const uint32_t *__restrict__ xv, const float *__restrict__ xs,
             float *__restrict__ ybuf, unsigned int *__restrict__ counters, int8_t *__restrict__ yv,
             float *__restrict__ ys, Key4 key, const uint64_t *__restrict__ tables) {
    extern __shared__ uint8_t smem_raw8[];
    Gemv8Smem<STAGES> &sm = *reinterpret_cast<Gemv8Smem<STAGES> *>((reinterpret_cast<uintptr_t>(smem_raw8) + 1023u) & ~(uintptr_t)1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t hb = cols >> 6, nitems = rows_local / kG8Rows;
    constexpr int kBPC = Gemv8Stage::kBlocksPerChunk;
    const uint32_t nchunks128 = (uint32_t)(cols >> 7);                    // 128-byte chunks per row
    const uint32_t steps = (nchunks128 + kG8Chunks - 1) / kG8Chunks;      // stages per work item

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1 + 32);              // TMA issuer (posts the tx bytes) + the 32 unit lanes
            mbar_init(&sm.empty[s], kG8Consumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kG8Consumers / 32) {
        // ------------------------------- TMA issuer -------------------------------
        if (lane == 0) {
            tma_prefetch_descriptor(&tmap);
            const uint64_t policy = policy_evict_first();
            uint32_t it = 0;
            for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
                for (uint32_t c = 0; c < steps; ++c, ++it) {
                    const int s = it % STAGES;
                    const uint32_t live = min((uint32_t)kG8Chunks, nchunks128 - c * kG8Chunks);
                    mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&sm.full[s], live * kG8ChunkBytes);
                    for (uint32_t j = 0; j < live; ++j)
                        tma_load_2d(sm.stage[s].rows[j], &tmap, (int)((c * kG8Chunks + j) * 128), (int)(item * kG8Rows),
                                    &sm.full[s], policy);
                }
            }
        }
    } else if (warp == kG8Consumers / 32 + 1) {
        // ------------------------------- x-unit warp -------------------------------
        // lane owns units lane + 32j (j = 0..4*kBPC/2-1) of a stage = (block 4j + lane/8, AVX lane lane%8); the raw operands
        // of a stage are requested one stage ahead, before the warp blocks on the slot (see k_m4_mvm_tma2)
        constexpr int kUnitsPerLane = kG8Chunks * kBPC / 4;
        uint32_t it = 0;
        const int l = lane & 7;
        uint32_t xa[kUnitsPerLane], xb[kUnitsPerLane];
        float sa[kUnitsPerLane], sb[kUnitsPerLane];
        auto prefetch = [&](uint64_t item, uint32_t c) {
            const float *su = scales + (item >> 1) * hb;
#pragma unroll
            for (int j = 0; j < kUnitsPerLane; ++j) {
                const uint64_t b = (uint64_t)c * (kBPC * kG8Chunks) + 4 * j + (lane >> 3);
                const bool ok = b < hb;
                xa[j] = ok ? __ldg(xv + b * 16 + l) : 0u;
                xb[j] = ok ? __ldg(xv + b * 16 + 8 + l) : 0u;
                sa[j] = ok ? __ldg(su + b) : 0.f;
                sb[j] = ok ? __ldg(xs + b) : 0.f;
            }
        };
        if (blockIdx.x < nitems) prefetch(blockIdx.x, 0);
        for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            for (uint32_t c = 0; c < steps; ++c, ++it) {
                const int s = it % STAGES;
                uint4 u[kUnitsPerLane];
#pragma unroll
                for (int j = 0; j < kUnitsPerLane; ++j) {
                    u[j].x = xa[j];
                    u[j].y = xb[j];
                    u[j].z = __float_as_uint(__fmul_rn(__fmul_rn(sa[j], 1.0f / 127.0f), __fmul_rn(sb[j], 1.0f / 127.0f)));    // (CloverMatrix8.h:1042-1046)
                    u[j].w = 0u;
                }
                uint32_t nc = c + 1; uint64_t nitem = item;
                if (nc == steps) { nc = 0; nitem = item + gridDim.x; }
                if (nitem < nitems) prefetch(nitem, nc);
                mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);
#pragma unroll
                for (int j = 0; j < kUnitsPerLane; ++j) sm.stage[s].units[lane + 32 * j] = u[j];
                mbar_arrive(&sm.full[s]);
            }
        }
    } else {
        // ------------------------------- consumer warps ------------------------------
        const int ty = warp & 1, rin = lane >> 2, l = 4 * ty + (lane & 3);
        const int r = 8 * (warp >> 1) + rin;                                 // row of the work item
        // byte offsets of this thread's two words of block b (b = 0, 1) inside a swizzled 32 x 128 B chunk
        const uint32_t base = (uint32_t)r * 128u + 4u * (uint32_t)(lane & 3);
        const uint32_t oa0 = base + ((uint32_t)((0 + ty) ^ rin) << 4), ob0 = base + ((uint32_t)((2 + ty) ^ rin) << 4);
        const uint32_t oa1 = base + ((uint32_t)((4 + ty) ^ rin) << 4), ob1 = base + ((uint32_t)((6 + ty) ^ rin) << 4);
        uint32_t it = 0;
        for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            float acc = 0.f;
            for (uint32_t c = 0; c < steps; ++c, ++it) {
                const int s = it % STAGES;
                mbar_wait(&sm.full[s], (it / STAGES) & 1);
                const Gemv8Stage &st = sm.stage[s];
                const int live = (int)min((uint32_t)kG8Chunks, nchunks128 - c * kG8Chunks);
                auto chunk = [&](int j) {
                    const uint8_t *p = st.rows[j];
                    const uint32_t wa0 = *reinterpret_cast<const uint32_t *>(p + oa0), wb0 = *reinterpret_cast<const uint32_t *>(p + ob0);
                    const uint32_t wa1 = *reinterpret_cast<const uint32_t *>(p + oa1), wb1 = *reinterpret_cast<const uint32_t *>(p + ob1);
                    const uint4 u0 = st.units[(2 * j) * 8 + l], u1 = st.units[(2 * j + 1) * 8 + l];
                    int d0 = dp4a_ss((int)wa0, (int)u0.x, (int)kMagicBits);
                    d0 = dp4a_ss((int)wb0, (int)u0.y, d0);
                    int d1 = dp4a_ss((int)wa1, (int)u1.x, (int)kMagicBits);
                    d1 = dp4a_ss((int)wb1, (int)u1.y, d1);
                    acc = __fmaf_rn(__uint_as_float(u0.z), __fsub_rn(__int_as_float(d0), 12582912.0f), acc);   // (:1093-1094)
                    acc = __fmaf_rn(__uint_as_float(u1.z), __fsub_rn(__int_as_float(d1), 12582912.0f), acc);
                };
                if (live == kG8Chunks) {
#pragma unroll
                    for (int j = 0; j < kG8Chunks; ++j) chunk(j);
                } else {
                    for (int j = 0; j < live; ++j) chunk(j);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
            }
            // hadd tree ((a4+a0)+(a6+a2))+((a5+a1)+(a7+a3)) (CloverBase.h:149-157); the 8 lanes of a row sit in two warps
            named_bar_sync(2, kG8Consumers);                                    // previous epilogue is done with part/ysm
            sm.part[r][l] = acc;
            named_bar_sync(2, kG8Consumers);
            const uint64_t rb = item >> 1, grb = rowblock0 + rb;
            if (tid < kG8Rows) {
                const float *a = sm.part[tid];
                const float y = __fadd_rn(__fadd_rn(__fadd_rn(a[4], a[0]), __fadd_rn(a[6], a[2])),
                                          __fadd_rn(__fadd_rn(a[5], a[1]), __fadd_rn(a[7], a[3])));
                ybuf[grb * 64 + (item & 1) * kG8Rows + tid] = y;
                __threadfence();
            }
            named_bar_sync(2, kG8Consumers);
            if (tid == 0) sm.ticket = yv ? atomicAdd(counters + rb, 1u) : 0u;
            named_bar_sync(2, kG8Consumers);
            if (sm.ticket == 1u) {                                              // both halves of the block are in ybuf
                if (tid < 64) {
                    __threadfence();
                    const float y = __ldcg(ybuf + grb * 64 + tid);
                    requantize_block<8, STOCH>(y, tid, grb, yv, ys, key, tables, sm.red_f, sm.red_q);
                    if (tid == 0) counters[rb] = 0u;                            // re-armed for the next launch
                }
            }
        }
    }
}

// =============================================================================================
// mixed-precision CloverMatrix4::mvm(V8,V8) (CloverMatrix4.h:1093-1441), round-2 kernel: thread = (row, chain PAIR)
// =============================================================================================
// k_m8_mvm_tma<.,.,4> gave every (row, AVX lane l) its own thread: two 16-bit shared loads + one 128-bit unit load and a
// three-instruction nibble expansion per halfword - 24 issue slots per (row, block, lane), LSU pipe 64-77 % busy, 71 % of
// HBM. Here a thread owns lanes l = 2p and 2p + 1 of one row:
//   * ONE 32-bit load per 16-byte half block (bytes 4p..4p+3 = the halfwords of lanes 2p and 2p+1), two PRMT sort the
//     four halfwords by lane: w_l = [A.b0 A.b1 B.b0 B.b1] = elements (e0|e1) (e2|e3) (e32|e33) (e34|e35) of lane l;
//   * the high nibbles (w & 0xF0F0F0F0 = 16*q of e0, e2, e32, e34) and the low nibbles ((w << 4) & 0xF0F0F0F0 = 16*q of
//     e1, e3, e33, e35) meet x operands the unit warp has already sorted the same way (xe, xo): LOP3, SHL, LOP3, 2 DP4A;
//   * units are 8 bytes per (block, lane) - one 128-bit load serves both lanes of the thread - and the block's scale
//     product is one broadcast word.
// 10 issue slots per (row, block, lane) instead of 24; 128 consumer threads per 32-row work item, KC 128-byte chunks per
// stage, so that more CTAs (smaller rings) share an SM. The hadd tree of a row lives in four adjacent lanes (shuffles).
// Arithmetic, chain order and the re-quantizer are those of k_m8_mvm<4>: results are bit-identical.
constexpr int kMxRows = 32;
constexpr int kMxConsumers = 128;
constexpr int kMxThreads = kMxConsumers + 64;
constexpr int kMxChunkBytes = kMxRows * 128;

template <int KC>
struct __align__(1024) MixStage {
    uint8_t rows[KC][kMxChunkBytes];          // chunk c: rows 0..31 x 128 B (256 nibbles = 4 blocks), SWIZZLE_128B
    uint2 units[KC * 4 * 8];                  // unit (block, l): x = xe (x0 x2 x32 x34 of the lane), y = xo (x1 x3 x33 x35)
    float prod[KC * 4];                       // (su * 1/7) * (sv * 1/127) of the block
};
template <int STAGES, int KC>
struct MixSmem {
    MixStage<KC> stage[STAGES];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    float red_f[2];
    int red_q[64];
    unsigned int ticket;
};

template <bool STOCH, int STAGES, int KC, int PER_SM>
__global__ void __launch_bounds__(kMxThreads, PER_SM)
k_m4v8_mvm_tma(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ scales, uint64_t rows_local,
               uint64_t cols, uint64_t rowblock0, const uint32_t *__restrict__ xv, const float *__restrict__ xs,
               float *__restrict__ ybuf, unsigned int *__restrict__ counters, int8_t *__restrict__ yv,
               float *__restrict__ ys, Key4 key, const uint64_t *__restrict__ tables) {
    extern __shared__ uint8_t smem_rawx[];
    MixSmem<STAGES, KC> &sm = *reinterpret_cast<MixSmem<STAGES, KC> *>((reinterpret_cast<uintptr_t>(smem_rawx) + 1023u) & ~(uintptr_t)1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t hb = cols >> 6, nitems = rows_local / kMxRows;
    const uint32_t nchunks128 = (uint32_t)((cols + 255) >> 8);            // the last chunk may be half outside the row: TMA zero-fills it
    const uint32_t steps = (nchunks128 + KC - 1) / KC;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1 + 32);              // TMA issuer (posts the tx bytes) + the 32 unit lanes
            mbar_init(&sm.empty[s], kMxConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kMxConsumers / 32) {
        // ------------------------------- TMA issuer -------------------------------
        if (lane == 0) {
            tma_prefetch_descriptor(&tmap);
            const uint64_t policy = policy_evict_first();
            uint32_t it = 0;
            for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
                for (uint32_t c = 0; c < steps; ++c, ++it) {
                    const int s = it % STAGES;
                    const uint32_t live = min((uint32_t)KC, nchunks128 - c * KC);
                    mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&sm.full[s], live * kMxChunkBytes);
                    for (uint32_t j = 0; j < live; ++j)
                        tma_load_2d(sm.stage[s].rows[j], &tmap, (int)((c * KC + j) * 128), (int)(item * kMxRows),
                                    &sm.full[s], policy);
                }
            }
        }
    } else if (warp == kMxConsumers / 32 + 1) {
        // ------------------------------- x-unit warp -------------------------------
        // lane owns units lane + 32j (j = 0..KC-1) of a stage = (block 4j + lane/8, AVX lane lane%8); the raw operands of a
        // stage are requested one stage ahead, before the warp blocks on the slot (see k_m4_mvm_tma2)
        uint32_t it = 0;
        const int l = lane & 7;
        uint32_t xa[KC], xb[KC];
        float sa[KC], sb[KC];
        auto prefetch = [&](uint64_t item, uint32_t c) {
            const float *su = scales + (item >> 1) * hb;
#pragma unroll
            for (int j = 0; j < KC; ++j) {
                const uint64_t b = (uint64_t)c * (4 * KC) + 4 * j + (lane >> 3);
                const bool ok = b < hb;
                xa[j] = ok ? __ldg(xv + b * 16 + l) : 0u;            // elements 4l .. 4l+3
                xb[j] = ok ? __ldg(xv + b * 16 + 8 + l) : 0u;        // elements 32+4l .. 32+4l+3
                sa[j] = ok ? __ldg(su + b) : 0.f;
                sb[j] = ok ? __ldg(xs + b) : 0.f;
            }
        };
        if (blockIdx.x < nitems) prefetch(blockIdx.x, 0);
        for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            for (uint32_t c = 0; c < steps; ++c, ++it) {
                const int s = it % STAGES;
                uint2 u[KC];
                float pr[KC];
#pragma unroll
                for (int j = 0; j < KC; ++j) {
                    u[j].x = __byte_perm(xa[j], xb[j], 0x6420);
                    u[j].y = __byte_perm(xa[j], xb[j], 0x7531);
                    pr[j] = __fmul_rn(__fmul_rn(sa[j], 1.0f / 7.0f), __fmul_rn(sb[j], 1.0f / 127.0f));
                }
                uint32_t nc = c + 1; uint64_t nitem = item;
                if (nc == steps) { nc = 0; nitem = item + gridDim.x; }
                if (nitem < nitems) prefetch(nitem, nc);
                mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);
#pragma unroll
                for (int j = 0; j < KC; ++j) {
                    sm.stage[s].units[lane + 32 * j] = u[j];
                    if (l == 0) sm.stage[s].prod[4 * j + (lane >> 3)] = pr[j];
                }
                mbar_arrive(&sm.full[s]);
            }
        }
    } else {
        // ------------------------------- consumer warps ------------------------------
        // warp = 8 rows, lane = (row rin of the eight, pair p): 16-byte chunk c of a row sits at chunk position c ^ (row & 7),
        // so the 32 words of a warp-wide load fall into 32 different banks
        const int rin = lane >> 2, p = lane & 3;
        const int r = 8 * warp + rin;
        const uint32_t base = (uint32_t)r * 128u + 4u * (uint32_t)p;
        uint32_t off[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) off[c] = base + ((uint32_t)(c ^ rin) << 4);
        uint32_t it = 0;
        for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            float acc0 = 0.f, acc1 = 0.f;
            for (uint32_t c = 0; c < steps; ++c, ++it) {
                const int s = it % STAGES;
                mbar_wait(&sm.full[s], (it / STAGES) & 1);
                // explicit shared-space loads with immediate offsets (a generic pointer into the aligned struct compiles to LD.E)
                const uint32_t sbase = smem_u32(&sm.stage[s]);
                uint32_t ra[8];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) ra[cc] = sbase + off[cc];
                const uint32_t ubase = sbase + 16u * (uint32_t)p;
                const int live = (int)min((uint32_t)KC, nchunks128 - c * KC);
#define CLOVER_MIX_BLOCK(J, B)                                                                                           \
    {                                                                                                                     \
        uint32_t wa, wb, pr_;                                                                                             \
        uint4 u;                                                                                                          \
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(wa) : "r"(ra[2 * (B)]), "n"((J) * kMxChunkBytes));              \
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(wb) : "r"(ra[2 * (B) + 1]), "n"((J) * kMxChunkBytes));          \
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)           \
                     : "r"(ubase), "n"(KC * kMxChunkBytes + (4 * (J) + (B)) * 64));                                       \
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(pr_) : "r"(sbase), "n"(KC * kMxChunkBytes + KC * 256 + (4 * (J) + (B)) * 4)); \
        const float pr = __uint_as_float(pr_);                                                                            \
        const uint32_t w0 = __byte_perm(wa, wb, 0x5410), w1 = __byte_perm(wa, wb, 0x7632);                                \
        /* d = magic + 16 S; as_float(d) / 16 - magic / 16 IS float(S), the reference's srai 4 + cvt (:1196) */           \
        int d0 = dp4a_ss((int)(w0 & 0xF0F0F0F0u), (int)u.x, (int)kMagicBits);                                             \
        d0 = dp4a_ss((int)((w0 << 4) & 0xF0F0F0F0u), (int)u.y, d0);                                                       \
        int d1 = dp4a_ss((int)(w1 & 0xF0F0F0F0u), (int)u.z, (int)kMagicBits);                                             \
        d1 = dp4a_ss((int)((w1 << 4) & 0xF0F0F0F0u), (int)u.w, d1);                                                       \
        acc0 = __fmaf_rn(pr, __fmaf_rn(__int_as_float(d0), 0.0625f, -786432.0f), acc0);                                   \
        acc1 = __fmaf_rn(pr, __fmaf_rn(__int_as_float(d1), 0.0625f, -786432.0f), acc1);                                   \
    }
#define CLOVER_MIX_CHUNK(J) { CLOVER_MIX_BLOCK(J, 0) CLOVER_MIX_BLOCK(J, 1) CLOVER_MIX_BLOCK(J, 2) CLOVER_MIX_BLOCK(J, 3) }
                if (live == KC) {
                    CLOVER_MIX_CHUNK(0) CLOVER_MIX_CHUNK(1) CLOVER_MIX_CHUNK(2) CLOVER_MIX_CHUNK(3)
                    if (KC == 8) { CLOVER_MIX_CHUNK(4) CLOVER_MIX_CHUNK(5) CLOVER_MIX_CHUNK(6) CLOVER_MIX_CHUNK(7) }
                } else {
                    // ragged last stage: same loads with a run-time chunk offset
                    for (int j = 0; j < live; ++j) {
                        uint32_t rj[8];
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc) rj[cc] = ra[cc] + (uint32_t)j * kMxChunkBytes;
                        const uint32_t uj = ubase + (uint32_t)j * 256u, pj = sbase + (uint32_t)j * 16u;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            uint32_t wa, wb, pr_;
                            uint4 u;
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa) : "r"(rj[2 * b]));
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb) : "r"(rj[2 * b + 1]));
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                         : "r"(uj + 64u * (uint32_t)b), "n"(KC * kMxChunkBytes));
                            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(pr_) : "r"(pj + 4u * (uint32_t)b), "n"(KC * kMxChunkBytes + KC * 256));
                            const float pr = __uint_as_float(pr_);
                            const uint32_t w0 = __byte_perm(wa, wb, 0x5410), w1 = __byte_perm(wa, wb, 0x7632);
                            int d0 = dp4a_ss((int)(w0 & 0xF0F0F0F0u), (int)u.x, (int)kMagicBits);
                            d0 = dp4a_ss((int)((w0 << 4) & 0xF0F0F0F0u), (int)u.y, d0);
                            int d1 = dp4a_ss((int)(w1 & 0xF0F0F0F0u), (int)u.z, (int)kMagicBits);
                            d1 = dp4a_ss((int)((w1 << 4) & 0xF0F0F0F0u), (int)u.w, d1);
                            acc0 = __fmaf_rn(pr, __fmaf_rn(__int_as_float(d0), 0.0625f, -786432.0f), acc0);
                            acc1 = __fmaf_rn(pr, __fmaf_rn(__int_as_float(d1), 0.0625f, -786432.0f), acc1);
                        }
                    }
                }
#undef CLOVER_MIX_CHUNK
#undef CLOVER_MIX_BLOCK
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
            }
            // hadd tree ((a4+a0)+(a6+a2))+((a5+a1)+(a7+a3)) (CloverBase.h:149-157): lane p holds a_2p, a_2p+1
            acc0 = __fadd_rn(acc0, __shfl_xor_sync(0xFFFFFFFFu, acc0, 2));      // p = 0,2: a4+a0   p = 1,3: a6+a2
            acc1 = __fadd_rn(acc1, __shfl_xor_sync(0xFFFFFFFFu, acc1, 2));      //          a5+a1             a7+a3
            acc0 = __fadd_rn(acc0, __shfl_xor_sync(0xFFFFFFFFu, acc0, 1));
            acc1 = __fadd_rn(acc1, __shfl_xor_sync(0xFFFFFFFFu, acc1, 1));
            const uint64_t rb = item >> 1, grb = rowblock0 + rb;
            if (p == 0) {
                ybuf[grb * 64 + (item & 1) * kMxRows + r] = __fadd_rn(acc0, acc1);
                __threadfence();
            }
            named_bar_sync(2, kMxConsumers);
            if (tid == 0) sm.ticket = yv ? atomicAdd(counters + rb, 1u) : 0u;
            named_bar_sync(2, kMxConsumers);
            if (sm.ticket == 1u) {                                              // both halves of the block are in ybuf
                if (tid < 64) {
                    __threadfence();
                    const float y = __ldcg(ybuf + grb * 64 + tid);
                    requantize_block<8, STOCH>(y, tid, grb, yv, ys, key, tables, sm.red_f, sm.red_q);
                    if (tid == 0) counters[rb] = 0u;                            // re-armed for the next launch
                }
            }
            named_bar_sync(2, kMxConsumers);                                    // ticket / red_* are free again
        }
    }
}

// =============================================================================================
// mvm(V4,V4), second-generation pipeline (BASELINE C3 default): the arithmetic of k_m4_mvm_tma in the layout of
// k_m8_mvm_tma - swizzled 32-row x 128-byte TMA boxes (no shared-memory bank conflicts; ncu counted 17 M conflicts
// per launch in the dense-box kernel), 32-row work items with a last-finisher re-quantize (2048 items on 148 SMs at
// 65536 rows; a 2-GPU shard still has 1024), one x-unit per (block, lane).
//   warp 16 (one lane)  eight cp.async.bulk.tensor.2d boxes (32 rows x 128 B = 256 nibbles = 4 blocks) per stage
//   warp 17             XUnits of the stage's 32 blocks
//   warps 0-7           thread = the reference's fp32 chain (accumulator a, AVX lane l) of TWO rows (r, r + 16: one
//                       XUnit load serves both); warp = 8 rows x 4 lanes with fixed (a, l >> 2): its two words per
//                       row and 128-byte chunk are 16-byte chunks 2a + (l>>2) and 2a + 4 + (l>>2), XOR-swizzled by
//                       (row & 7) -> 32 lanes, 32 banks.
// =============================================================================================
constexpr int kG4Rows = 32;
constexpr int kG4Chunks = 8;                 // 128-byte chunks per stage = 32 blocks of 64 columns
constexpr int kG4Consumers = 256;
constexpr int kG4Threads = kG4Consumers + 64;
constexpr int kG4ChunkBytes = kG4Rows * 128;

struct __align__(1024) Gemv4Stage {
    uint8_t rows[kG4Chunks][kG4ChunkBytes];
    XUnit units[kG4Chunks * 4 * 8];          // unit (block, l)
};
// STAGES = 5: one CTA per SM (180 KiB ring). STAGES = 3: 111 KiB, TWO CTAs per SM - the grid is then 2 x SMs and the
// hardware shares each SM between two work items, which removes the whole-round quantisation of a persistent grid: a
// shard of 256 or 512 items (8 / 4 GPUs) no longer leaves 14 % of the SM-rounds empty, all items simply stream
// concurrently at the HBM rate (216 KiB in flight per SM instead of 144).
template <int STAGES>
struct Gemv4Smem {
    Gemv4Stage stage[STAGES];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    float part[kG4Rows][16];
    float red_f[2];
    int red_q[64];
    unsigned int ticket;
    int peers_started;                       // stamped exchange: every peer has started this call
};

template <bool STOCH, int STAGES, int XCHG>
__global__ void __launch_bounds__(kG4Threads, STAGES == 5 ? 1 : 2)
k_m4_mvm_tma2(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ scales, uint64_t rows_local,
              uint64_t cols, uint64_t rowblock0, const uint32_t *__restrict__ xv, const float *__restrict__ xs,
              float *__restrict__ ybuf, unsigned int *__restrict__ counters, int8_t *__restrict__ yv,
              float *__restrict__ ys, Key4 key, const uint64_t *__restrict__ tables, const __grid_constant__ PeerOut peers) {
    extern __shared__ uint8_t smem_raw4[];
    Gemv4Smem<STAGES> &sm = *reinterpret_cast<Gemv4Smem<STAGES> *>((reinterpret_cast<uintptr_t>(smem_raw4) + 1023u) & ~(uintptr_t)1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t hb = cols >> 6, nitems = rows_local / kG4Rows;
    const uint32_t nchunks128 = (uint32_t)((cols + 255) >> 8);            // 128-byte chunks per row (256 nibbles each); the last one may be half
                                                                          // outside the row (cols = 128 mod 256): TMA zero-fills it, its units are zero
    const uint32_t steps = (nchunks128 + kG4Chunks - 1) / kG4Chunks;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1 + 32);
            mbar_init(&sm.empty[s], kG4Consumers / 32);
        }
        sm.peers_started = 0;
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kG4Consumers / 32) {
        // ------------------------------- TMA issuer -------------------------------
        if (lane == 0) {
            tma_prefetch_descriptor(&tmap);
            const uint64_t policy = policy_evict_first();
            uint32_t it = 0;
            for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
                for (uint32_t c = 0; c < steps; ++c, ++it) {
                    const int s = it % STAGES;
                    const uint32_t live = min((uint32_t)kG4Chunks, nchunks128 - c * kG4Chunks);
                    mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&sm.full[s], live * kG4ChunkBytes);
                    for (uint32_t j = 0; j < live; ++j)
                        tma_load_2d(sm.stage[s].rows[j], &tmap, (int)((c * kG4Chunks + j) * 128), (int)(item * kG4Rows),
                                    &sm.full[s], policy);
                }
            }
        } else if (XCHG == kXchgStamped && peers.world > 1) {
            stamped_flow_control(peers, &sm.peers_started);   // the issuer warp's idle lanes
        }
    } else if (warp == kG4Consumers / 32 + 1) {
        // ------------------------------- x-unit warp -------------------------------
        // lane owns units lane + 32j (j = 0..7) of a stage = (block 4j + lane/8, AVX lane lane%8)
        uint32_t it = 0;
        const int l = lane & 7;
        // The raw operands of a stage (x word, matrix-tile scale, x scale per unit) are requested ONE STAGE AHEAD, before this
        // warp blocks on the slot: their L2 latency (~1 us while the matrix streams at the HBM rate) never adds to the
        // consumers' stage time. Without it a unit warp that once falls behind the TMA ring (a late start is enough)
        // stays behind for the whole kernel, because every stage then costs (load latency +
        // consume) instead of max(load latency, consume).
        uint32_t w[8];
        float sa[8], sb[8];
        auto prefetch = [&](uint64_t item, uint32_t c) {
            const float *su = scales + (item >> 1) * hb;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint64_t b = (uint64_t)c * (4 * kG4Chunks) + 4 * j + (lane >> 3);
                const bool ok = b < hb;
                w[j] = ok ? __ldg(xv + b * 8 + l) : 0u;
                sa[j] = ok ? __ldg(su + b) : 0.f;
                sb[j] = ok ? __ldg(xs + b) : 0.f;
            }
        };
        if (blockIdx.x < nitems) prefetch(blockIdx.x, 0);
        for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            for (uint32_t c = 0; c < steps; ++c, ++it) {
                const int s = it % STAGES;
                XUnit u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int xh = sext_nibbles((w[j] >> 4) & 0x0F0F0F0Fu);
                    const int xl = sext_nibbles(w[j] & 0x0F0F0F0Fu);
                    u[j].xh = xh;
                    u[j].xl16 = (int)((w[j] << 4) & 0xF0F0F0F0u);
                    u[j].cneg = -(786432.0f + 8.0f * (float)(dp4a_ss(xh, 0x01010101, 0) + dp4a_ss(xl, 0x01010101, 0)));
                    u[j].prod = __fmul_rn(__fmul_rn(sa[j], 1.0f / 49.0f), sb[j]);                 // (:834-837)
                }
                uint32_t nc = c + 1; uint64_t nitem = item;
                if (nc == steps) { nc = 0; nitem = item + gridDim.x; }
                if (nitem < nitems) prefetch(nitem, nc);
                mbar_wait(&sm.empty[s], ((it / STAGES) & 1) ^ 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) sm.stage[s].units[lane + 32 * j] = u[j];
                mbar_arrive(&sm.full[s]);
            }
        }
    } else {
        // ------------------------------- consumer warps ------------------------------
        // warp = (row group rg, accumulator a, lane type ty); lane = (row rin of the group, l & 3); rows r and r + 16
        const int rg = warp >> 2, a = (warp >> 1) & 1, ty = warp & 1;
        const int rin = lane >> 2, l = 4 * ty + (lane & 3);
        const int r = 8 * rg + rin;
        const uint32_t base = (uint32_t)r * 128u + 4u * (uint32_t)(lane & 3);
        // this chain's two blocks of a 128-byte chunk: b = a and b = a + 2 -> 16-byte chunks 2b + ty, swizzled by (row & 7)
        const uint32_t o0 = base + ((uint32_t)((2 * a + ty) ^ rin) << 4), o1 = base + ((uint32_t)((2 * a + 4 + ty) ^ rin) << 4);
        uint32_t it = 0;
        for (uint64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            float acc = 0.f, acc2 = 0.f;
            for (uint32_t c = 0; c < steps; ++c, ++it) {
                const int s = it % STAGES;
                mbar_wait(&sm.full[s], (it / STAGES) & 1);
                const Gemv4Stage &st = sm.stage[s];
                const int live = (int)min((uint32_t)kG4Chunks, nchunks128 - c * kG4Chunks);
                auto chunk = [&](int j) {
                    const uint8_t *p = st.rows[j];
                    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(p + o0), w1 = *reinterpret_cast<const uint32_t *>(p + o1);
                    const uint32_t v0 = *reinterpret_cast<const uint32_t *>(p + o0 + 16 * 128), v1 = *reinterpret_cast<const uint32_t *>(p + o1 + 16 * 128);
                    const XUnit u0 = st.units[(4 * j + a) * 8 + l], u1 = st.units[(4 * j + a + 2) * 8 + l];
                    int s0 = dp4a_us(xor_and(w0, 0x88888888u, 0xF0F0F0F0u), u0.xh, (int)kMagicBits);
                    s0 = dp4a_us(xor_and(w0, 0x88888888u, 0x0F0F0F0Fu), u0.xl16, s0);
                    int t0 = dp4a_us(xor_and(v0, 0x88888888u, 0xF0F0F0F0u), u0.xh, (int)kMagicBits);
                    t0 = dp4a_us(xor_and(v0, 0x88888888u, 0x0F0F0F0Fu), u0.xl16, t0);
                    int s1 = dp4a_us(xor_and(w1, 0x88888888u, 0xF0F0F0F0u), u1.xh, (int)kMagicBits);
                    s1 = dp4a_us(xor_and(w1, 0x88888888u, 0x0F0F0F0Fu), u1.xl16, s1);
                    int t1 = dp4a_us(xor_and(v1, 0x88888888u, 0xF0F0F0F0u), u1.xh, (int)kMagicBits);
                    t1 = dp4a_us(xor_and(v1, 0x88888888u, 0x0F0F0F0Fu), u1.xl16, t1);
                    acc = __fmaf_rn(u0.prod, __fmaf_rn(__int_as_float(s0), 0.0625f, u0.cneg), acc);      // (:896-897), block order
                    acc2 = __fmaf_rn(u0.prod, __fmaf_rn(__int_as_float(t0), 0.0625f, u0.cneg), acc2);
                    acc = __fmaf_rn(u1.prod, __fmaf_rn(__int_as_float(s1), 0.0625f, u1.cneg), acc);
                    acc2 = __fmaf_rn(u1.prod, __fmaf_rn(__int_as_float(t1), 0.0625f, u1.cneg), acc2);
                };
                if (live == kG4Chunks) {
#pragma unroll
                    for (int j = 0; j < kG4Chunks; ++j) chunk(j);
                } else {
                    for (int j = 0; j < live; ++j) chunk(j);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
            }
            named_bar_sync(2, kG4Consumers);                                    // previous epilogue is done with part
            sm.part[r][8 * a + l] = acc;
            sm.part[r + 16][8 * a + l] = acc2;
            named_bar_sync(2, kG4Consumers);
            const uint64_t rb = item >> 1, grb = rowblock0 + rb;
            if (tid < kG4Rows) {
                const float *pa = sm.part[tid];
                float sl[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) sl[k] = __fadd_rn(pa[k], pa[8 + k]);                   // acc_1 + acc_2 (:902)
                const float y = __fadd_rn(__fadd_rn(__fadd_rn(sl[4], sl[0]), __fadd_rn(sl[6], sl[2])),
                                          __fadd_rn(__fadd_rn(sl[5], sl[1]), __fadd_rn(sl[7], sl[3])));   // hadd tree (:903-907)
                ybuf[grb * 64 + (item & 1) * kG4Rows + tid] = y;
                __threadfence();
            }
            named_bar_sync(2, kG4Consumers);
            if (tid == 0) sm.ticket = yv ? atomicAdd(counters + rb, 1u) : 0u;
            named_bar_sync(2, kG4Consumers);
            if (sm.ticket == 1u) {                                              // both halves of the block are in ybuf
                if (tid < 64) {
                    __threadfence();
                    const float y = __ldcg(ybuf + grb * 64 + tid);
                    if (XCHG == kXchgStamped && peers.world > 1) { while (*(volatile int *)&sm.peers_started == 0) { } }
                    requantize_block<4, STOCH>(y, tid, grb, yv, ys, key, tables, sm.red_f, sm.red_q,
                                               peers.world > 1 ? &peers : nullptr);
                    if (tid == 0) counters[rb] = 0u;
                }
            }
        }
        if (XCHG != kXchgStamped && peers.world > 1) {
            if (tid < 64) __threadfence_system();                // this CTA's peer stores are visible system-wide ...
            named_bar_sync(2, kG4Consumers);                     // ... before its ticket is taken
            if (warp == 0) peer_signal_and_wait(peers);        // all 32 lanes of consumer warp 0
        }
    }
}

// re-quantize a full fp32 vector like the mvm tail (used after the multi-GPU exchange)
template <int BITS, bool STOCH>
__global__ void __launch_bounds__(64)
k_requantize_mvm(const float *__restrict__ y32, uint64_t nblocks, int8_t *__restrict__ yv, float *__restrict__ ys,
                 Key4 key, const uint64_t *__restrict__ tables) {
    __shared__ float red_f[2];
    __shared__ int red_q[64];
    for (uint64_t rb = blockIdx.x; rb < nblocks; rb += gridDim.x) {
        requantize_block<BITS, STOCH>(y32[rb * 64 + threadIdx.x], threadIdx.x, rb, yv, ys, key, tables, red_f, red_q);
        asm volatile("bar.sync 1, 64;");
    }
}

// scratch of the pipelined GEMVs: fp32 row results (when the caller passes no y32) and one zero-initialised counter per
// row block, re-armed by the kernel itself. Keyed by (device, stream): calls on different streams never share counters.
static int mvm_scratch(cudaStream_t stream, uint64_t nrb_local, uint64_t nrb_global_end, float **ybuf, unsigned int **counters) {
    void *p = nullptr;
    if (ybuf) {
        int rc = stream_scratch(kScratchMvmY, stream, nrb_global_end * 64 * sizeof(float), 0, &p);
        if (rc != CLOVER_OK) return rc;
        *ybuf = static_cast<float *>(p);
    }
    // counters: every counter of a block is zero between launches, so the whole (possibly larger) new block is zeroed
    int rc = stream_scratch(kScratchMvmCounters, stream, nrb_local * sizeof(unsigned int), ~(size_t)0, &p);
    if (rc != CLOVER_OK) return rc;
    *counters = static_cast<unsigned int *>(p);
    return CLOVER_OK;
}

// the TMA-ring kernel of CloverMatrix8::mvm
static int launch_mvm8_tma(const int8_t *values, const float *scales, uint64_t rows_local, uint64_t cols, uint64_t row0,
                           const uint32_t *x32, const float *xs, float *y32, int8_t *yv, float *ys, bool stoch, Key4 key,
                           const uint64_t *tables, cudaStream_t stream) {
    const uint64_t nrb = rows_local >> 6;
    // fp32 row results pass through global memory (the caller's y32, else a grow-only per-stream scratch) and a
    // zero-initialised counter per row block, re-armed by the kernel itself.
    float *ybuf = y32;
    unsigned int *counters = nullptr;
    int rc = mvm_scratch(stream, nrb, (row0 >> 6) + nrb, y32 ? nullptr : &ybuf, &counters);
    if (rc != CLOVER_OK) return rc;
    // two CTAs per SM with 3-stage rings, as for the 4-bit kernel, whenever there is more than one round of work items
    // (tools/gemv_shapes.py 60 8: 32768^2 164 -> 160 us, 16384 x 32768 93 -> 83, 8192 x 32768 48 -> 42, 16384 x 4096
    // 20.7 -> 16.4; a single round - 4096 x 32768 - prefers the deeper ring: 23.0 vs 24.8). CLOVER_GEMV_IMPL=items32 /
    // items32x2 force either.
    const char *impl8 = getenv("CLOVER_GEMV_IMPL");
    const bool x2 = impl8 ? !strcmp(impl8, "items32x2") : rows_local / kG8Rows > (uint64_t)sm_count();
    const int smem = (int)(x2 ? sizeof(Gemv8Smem<3>) : sizeof(Gemv8Smem<5>)) + 1024;
    // function attributes and occupancy belong to a device's context: remembered per device (clover_set_device may switch)
    static bool attr_set8[kMaxDevices][2][2] = {};
    const int dev = current_device();
    if (dev < 0) { set_error("device index out of range"); return CLOVER_ERR_INVALID; }
    auto kern = x2 ? (stoch ? k_m8_mvm_tma<true, 3> : k_m8_mvm_tma<false, 3>) : (stoch ? k_m8_mvm_tma<true, 5> : k_m8_mvm_tma<false, 5>);
    if (!attr_set8[dev][x2][stoch]) {
        CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set8[dev][x2][stoch] = true;
    }
    CUtensorMap tmap;
    rc = make_tensor_map_u8_2d_sw128(&tmap, values, rows_local, cols, kG8Rows);
    if (rc != CLOVER_OK) return rc;
    const uint64_t nitems = rows_local / kG8Rows;
    uint64_t slots = (uint64_t)sm_count();
    if (x2) {
        static int per_sm8[kMaxDevices][2] = {};
        if (!per_sm8[dev][stoch]) CLOVER_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm8[dev][stoch], kern, kG8Threads, smem));
        slots *= (uint64_t)std::max(1, std::min(per_sm8[dev][stoch], 2));
    }
    const unsigned pgrid = (unsigned)(nitems < slots ? nitems : slots);
    kern<<<pgrid, kG8Threads, smem, stream>>>(tmap, scales, rows_local, cols, row0 >> 6, x32, xs, ybuf, counters, yv, ys,
                                              key, tables);
    return CLOVER_OK;
}

// the mixed-precision TMA-ring kernel (k_m4v8_mvm_tma): ring geometry by shape
template <bool STOCH, int STAGES, int KC, int PER_SM>
static int launch_mix_variant(const CUtensorMap &tmap, const float *scales, uint64_t rows_local, uint64_t cols, uint64_t row0,
                              const uint32_t *x32, const float *xs, float *ybuf, unsigned int *counters, int8_t *yv, float *ys,
                              Key4 key, const uint64_t *tables, cudaStream_t stream) {
    static bool attr_set[kMaxDevices] = {};
    static int per_sm[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0) { set_error("device index out of range"); return CLOVER_ERR_INVALID; }
    auto kern = k_m4v8_mvm_tma<STOCH, STAGES, KC, PER_SM>;
    const int smem = (int)sizeof(MixSmem<STAGES, KC>) + 1024;
    if (!attr_set[dev]) {
        CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CLOVER_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[dev], kern, kMxThreads, smem));
        attr_set[dev] = true;
    }
    const uint64_t nitems = rows_local / kMxRows;
    const uint64_t slots = (uint64_t)sm_count() * (uint64_t)std::max(1, std::min(per_sm[dev], PER_SM));
    const unsigned pgrid = (unsigned)(nitems < slots ? nitems : slots);
    kern<<<pgrid, kMxThreads, smem, stream>>>(tmap, scales, rows_local, cols, row0 >> 6, x32, xs, ybuf, counters, yv, ys, key, tables);
    return CLOVER_OK;
}

static int launch_mix_tma(const int8_t *values, const float *scales, uint64_t rows_local, uint64_t cols, uint64_t row0,
                          const uint32_t *x32, const float *xs, float *y32, int8_t *yv, float *ys, bool stoch, Key4 key,
                          const uint64_t *tables, cudaStream_t stream) {
    const uint64_t nrb = rows_local >> 6;
    float *ybuf = y32;
    unsigned int *counters = nullptr;
    int rc = mvm_scratch(stream, nrb, (row0 >> 6) + nrb, y32 ? nullptr : &ybuf, &counters);
    if (rc != CLOVER_OK) return rc;
    CUtensorMap tmap;
    rc = make_tensor_map_u8_2d_sw128(&tmap, values, rows_local, cols >> 1, kMxRows);
    if (rc != CLOVER_OK) return rc;
#define CLOVER_MIX(ST, KC, PS)                                                                                                         \
    (stoch ? launch_mix_variant<true, ST, KC, PS>(tmap, scales, rows_local, cols, row0, x32, xs, ybuf, counters, yv, ys, key, tables, stream) \
           : launch_mix_variant<false, ST, KC, PS>(tmap, scales, rows_local, cols, row0, x32, xs, ybuf, counters, yv, ys, key, tables, stream))
    // measured on B200 (tools/mix_sweep.py; us with four 52 KiB CTAs per SM / two 106 KiB CTAs per SM): 32768^2 88.1 / 91.1,
    // 16384^2 28.9 / 29.9, 32768 x 8192 28.9 / 33.0 - but 8192 x 32768 29.0 / 26.9 and 4096 x 32768 18.8 / 16.7: with fewer
    // work items than two per SM the deeper ring wins. CLOVER_GEMV_IMPL=ring4 / ring8 force either.
    const char *impl = getenv("CLOVER_GEMV_IMPL");
    const bool small_rings = impl && !strcmp(impl, "ring4") ? true : impl && !strcmp(impl, "ring8") ? false
                                                            : rows_local / kMxRows > 2 * (uint64_t)sm_count();
    return small_rings ? CLOVER_MIX(3, 4, 4) : CLOVER_MIX(3, 8, 2);
#undef CLOVER_MIX
}

template <int BITS>
static int launch_mvm(const int8_t *values, const float *scales, uint64_t rows_local, uint64_t cols, uint64_t row0,
                      const int8_t *xv, const float *xs, float *y32, int8_t *yv, float *ys, const uint64_t *key_host,
                      cudaStream_t stream, const PeerOut *peers = nullptr) {
    const uint64_t nrb = rows_local >> 6;
    if (nrb == 0 || cols == 0) return CLOVER_OK;
    const unsigned grid = (unsigned)nrb;
    Key4 key = {};
    const uint64_t *tables = nullptr;
    const bool stoch = key_host != nullptr && yv != nullptr;
    if (stoch) {
        tables = device_jump_tables();
        if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
        key = key_lanes(key_host);
    }
    const uint32_t *v32 = reinterpret_cast<const uint32_t *>(values);
    const uint32_t *x32 = reinterpret_cast<const uint32_t *>(xv);
    if (BITS == 4) {
        // CLOVER_GEMV_IMPL (read per call): simple = plain loads, ring64 = 64-row work items (k_m4_mvm_tma),
        // items32 = 32-row work items (k_m4_mvm_tma2), items32x2 = the same at two CTAs per SM.
        const char *impl_env = getenv("CLOVER_GEMV_IMPL");
        const bool force_simple = impl_env && !strcmp(impl_env, "simple");
        const bool force_items32 = impl_env && !strcmp(impl_env, "items32");
        const bool force_x2 = impl_env && !strcmp(impl_env, "items32x2");
        auto fill = [](uint64_t items, uint64_t sms) { return (double)items / (double)(((items + sms - 1) / sms) * sms); };
        // Default: 32-row items at TWO CTAs per SM (3-stage rings). Measured on B200 (tools/gemv_shapes.py, 65536 columns):
        // 65536 rows 317 us (ring64 335, items32 356), 32768 rows 169 (186 / 179), 16384 rows 88 (93 / 103), 8192 rows 43
        // (46 / 55) - the SM is shared by two work items, so a persistent grid no longer runs in whole rounds and 216 KiB
        // per SM are in flight. Short rows (< 16384 columns: 32768 x 8192 runs 31.4 us under ring64, 33.6 under x2) keep
        // the 64-row items, whose epilogue is paid half as often.
        const bool x2 = force_x2 || (!impl_env && cols >= 16384);
        const bool force_ring64 = !x2 && ((impl_env && !strcmp(impl_env, "ring64")) ||
                                  (!force_items32 && (cols < 16384 || fill(2 * nrb, (uint64_t)sm_count()) < 1.04 * fill(nrb, (uint64_t)sm_count()))));
        // bulk copies need 16 B aligned rows; anything else takes the plain-load kernel (same arithmetic)
        const bool simple = (force_simple || (reinterpret_cast<uintptr_t>(values) & 15u) != 0) && !peers;
        if (simple) {
            if (stoch) k_m4_mvm<true><<<grid, kMvmThreads, 0, stream>>>(v32, scales, rows_local, cols, row0 >> 6, x32, xs, y32, yv, ys, key, tables);
            else       k_m4_mvm<false><<<grid, kMvmThreads, 0, stream>>>(v32, scales, rows_local, cols, row0 >> 6, x32, xs, y32, yv, ys, key, tables);
        } else if (!force_ring64) {
            // default: 32-row work items, swizzled boxes (k_m4_mvm_tma2)
            float *ybuf = y32;
            unsigned int *counters = nullptr;
            int rc = mvm_scratch(stream, nrb, (row0 >> 6) + nrb, y32 ? nullptr : &ybuf, &counters);
            if (rc != CLOVER_OK) return rc;
            const int smem = (int)(x2 ? sizeof(Gemv4Smem<3>) : sizeof(Gemv4Smem<5>)) + 1024;
            static bool attr_set2[kMaxDevices][2][2][3] = {};   // per device: function attributes belong to a device's context
            const int dev = current_device();
            if (dev < 0) { set_error("device index out of range"); return CLOVER_ERR_INVALID; }
            // the stamped exchange is its own instantiation (the local / flag-synchronised kernel stays as it was)
            const int xchg = peers && peers->world > 1 ? peers->defer : kXchgFlags;
            using KernT = void (*)(const CUtensorMap, const float *, uint64_t, uint64_t, uint64_t, const uint32_t *, const float *, float *,
                                   unsigned int *, int8_t *, float *, Key4, const uint64_t *, const PeerOut);
            static const KernT kerns[2][2][2] = {     // [x2][stoch][stamped]
                {{k_m4_mvm_tma2<false, 5, kXchgFlags>, k_m4_mvm_tma2<false, 5, kXchgStamped>},
                 {k_m4_mvm_tma2<true, 5, kXchgFlags>, k_m4_mvm_tma2<true, 5, kXchgStamped>}},
                {{k_m4_mvm_tma2<false, 3, kXchgFlags>, k_m4_mvm_tma2<false, 3, kXchgStamped>},
                 {k_m4_mvm_tma2<true, 3, kXchgFlags>, k_m4_mvm_tma2<true, 3, kXchgStamped>}}};
            const KernT kern = kerns[x2][stoch][xchg == kXchgStamped];
            if (!attr_set2[dev][x2][stoch][xchg]) {
                CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                attr_set2[dev][x2][stoch][xchg] = true;
            }
            CUtensorMap tmap;
            rc = make_tensor_map_u8_2d_sw128(&tmap, values, rows_local, cols >> 1, kG4Rows);
            if (rc != CLOVER_OK) return rc;
            const uint64_t nitems = rows_local / kG4Rows;
            uint64_t slots = (uint64_t)sm_count();
            if (x2) {
                static int per_sm4[kMaxDevices][2][3] = {};   // asked once per template instance and device
                if (!per_sm4[dev][stoch][xchg]) CLOVER_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm4[dev][stoch][xchg], kern, kG4Threads, smem));
                slots *= (uint64_t)std::max(1, std::min(per_sm4[dev][stoch][xchg], 2));
            }
            const unsigned pgrid = (unsigned)(nitems < slots ? nitems : slots);
            kern<<<pgrid, kG4Threads, smem, stream>>>(tmap, scales, rows_local, cols, row0 >> 6, x32, xs, ybuf, counters, yv, ys,
                                                      key, tables, peers ? *peers : PeerOut());
        } else {
            const int smem = (int)sizeof(GemvSmem);
            static bool attr_set[kMaxDevices][2][3] = {};   // per template instance and device
            const int dev = current_device();
            if (dev < 0) { set_error("device index out of range"); return CLOVER_ERR_INVALID; }
            const int xchg = peers && peers->world > 1 ? peers->defer : kXchgFlags;
            using KernT = void (*)(const CUtensorMap, const float *, uint64_t, uint64_t, uint64_t, const uint32_t *, const float *, float *,
                                   int8_t *, float *, Key4, const uint64_t *, const PeerOut);
            static const KernT kerns[2][2] = {{k_m4_mvm_tma<false, kXchgFlags>, k_m4_mvm_tma<false, kXchgStamped>},
                                              {k_m4_mvm_tma<true, kXchgFlags>, k_m4_mvm_tma<true, kXchgStamped>}};
            const KernT kern = kerns[stoch][xchg == kXchgStamped];
            if (!attr_set[dev][stoch][xchg]) {
                CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                attr_set[dev][stoch][xchg] = true;
            }
            CUtensorMap tmap;
            int rc = make_tensor_map_u32_2d(&tmap, values, rows_local, cols >> 3, cols >> 1, 32, kKC * 8);
            if (rc != CLOVER_OK) return rc;
            const unsigned pgrid = (unsigned)(nrb < (uint64_t)sm_count() ? nrb : (uint64_t)sm_count());
            kern<<<pgrid, kGemvThreads, smem, stream>>>(tmap, scales, rows_local, cols, row0 >> 6, x32, xs, y32, yv, ys,
                                                        key, tables, peers ? *peers : PeerOut());
        }
    } else {
        static const bool force_simple8 = getenv("CLOVER_GEMV_IMPL") && !strcmp(getenv("CLOVER_GEMV_IMPL"), "simple");
        const bool simple = force_simple8 || (reinterpret_cast<uintptr_t>(values) & 15u) != 0;
        if (simple) {
            if (stoch) k_m8_mvm<8, true><<<grid, 512, 0, stream>>>(v32, scales, rows_local, cols, row0 >> 6, x32, xs, y32, yv, ys, key, tables);
            else       k_m8_mvm<8, false><<<grid, 512, 0, stream>>>(v32, scales, rows_local, cols, row0 >> 6, x32, xs, y32, yv, ys, key, tables);
        } else {
            int rc = launch_mvm8_tma(values, scales, rows_local, cols, row0, x32, xs, y32, yv, ys, stoch, key, tables, stream);
            if (rc != CLOVER_OK) return rc;
        }
    }
    count_launch();
    return launch_status("k_mvm");
}

}  // namespace clover

using namespace clover;

#define CLOVER_CHECK_MAT(rows, cols)                                                                   \
    CLOVER_REQUIRE((rows) % 128u == 0 && (cols) % 128u == 0, CLOVER_ERR_INVALID,                       \
                   "rows and cols must be multiples of 128 (include/CloverMatrix.h:48-50)")

extern "C" {

int clover_m4_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales,
                       uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(a && values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_MAT(rows, cols);
    return launch_mquantize<4>(a, rows, cols, values, scales, key_host, (cudaStream_t)stream);
}
int clover_m8_quantize(const float *a, uint64_t rows, uint64_t cols, int8_t *values, float *scales,
                       uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(a && values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_MAT(rows, cols);
    return launch_mquantize<8>(a, rows, cols, values, scales, key_host, (cudaStream_t)stream);
}

int clover_m4_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                  const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32,
                  uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(values && scales && xv && xs && yv && ys, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_MAT(rows, cols);
    int rc = launch_mvm<4>(values, scales, rows, cols, 0, xv, xs, y32, yv, ys, key_host, (cudaStream_t)stream);
    if (rc == CLOVER_OK && key_host) host_key_skip(key_host, 2 * (rows >> 6));   // two calls per 64 rows (:979,:999)
    return rc;
}
int clover_m8_mvm(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                  const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32,
                  uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(values && scales && xv && xs && yv && ys, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_MAT(rows, cols);
    int rc = launch_mvm<8>(values, scales, rows, cols, 0, xv, xs, y32, yv, ys, key_host, (cudaStream_t)stream);
    if (rc == CLOVER_OK && key_host) host_key_skip(key_host, 2 * (rows >> 6));
    return rc;
}

int clover_m4_mvm_v8(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                     const int8_t *xv, const float *xs, int8_t *yv, float *ys, float *y32,
                     uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(values && scales && xv && xs && yv && ys, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_MAT(rows, cols);
    const uint64_t nrb = rows >> 6;
    if (nrb == 0 || cols == 0) return CLOVER_OK;
    Key4 key = {};
    const uint64_t *tables = nullptr;
    if (key_host) {
        tables = device_jump_tables();
        if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
        key = key_lanes(key_host);
    }
    const uint32_t *v32 = reinterpret_cast<const uint32_t *>(values), *x32 = reinterpret_cast<const uint32_t *>(xv);
    cudaStream_t st = (cudaStream_t)stream;
    // CLOVER_GEMV_IMPL=simple (or an unaligned matrix): plain-load kernel; default: the TMA ring of the 8-bit GEMV on nibble rows
    const char *impl = getenv("CLOVER_GEMV_IMPL");
    if ((impl && !strcmp(impl, "simple")) || (reinterpret_cast<uintptr_t>(values) & 15u) != 0) {
        if (key_host) k_m8_mvm<4, true><<<(unsigned)nrb, 512, 0, st>>>(v32, scales, rows, cols, 0, x32, xs, y32, yv, ys, key, tables);
        else          k_m8_mvm<4, false><<<(unsigned)nrb, 512, 0, st>>>(v32, scales, rows, cols, 0, x32, xs, y32, yv, ys, key, tables);
    } else {
        const int rc2 = launch_mix_tma(values, scales, rows, cols, 0, x32, xs, y32, yv, ys, key_host != nullptr, key, tables, st);
        if (rc2 != CLOVER_OK) return rc2;
    }
    count_launch();
    const int rc = launch_status("k_m4_mvm_v8");
    if (rc == CLOVER_OK && key_host) host_key_skip(key_host, 2 * nrb);
    return rc;
}

int clover_m4_mvm_shard(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                        uint64_t row0, const int8_t *xv, const float *xs, float *y32_full,
                        int8_t *yv_full, float *ys_full, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(values_local && scales_local && xv && xs, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(y32_full || (yv_full && ys_full), CLOVER_ERR_INVALID, "no output requested");
    CLOVER_REQUIRE(rows_local % 64u == 0 && row0 % 64u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID,
                   "shards are whole 64-row blocks; cols a multiple of 128");
    // key_host (if any) is read at the GLOBAL row-block position and NOT advanced: the caller advances
    // it once by 2 * total_rows / 64 (clover_prng_skip), so every rank stays on the reference's stream.
    return launch_mvm<4>(values_local, scales_local, rows_local, cols, row0, xv, xs, y32_full,
                         yv_full, ys_full, key_host, (cudaStream_t)stream);
}

static int launch_shard_fused(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                              uint64_t row0, const int8_t *xv, const float *xs, int8_t *const *peer_yv_host,
                              float *const *peer_ys_host, uint32_t *const *peer_flags_host, unsigned int *ticket,
                              int world, int rank, uint32_t epoch, uint64_t *key_host, void *stream, int defer) {
    CLOVER_REQUIRE(values_local && scales_local && xv && xs && peer_yv_host && peer_ys_host && peer_flags_host && ticket,
                   CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, CLOVER_ERR_INVALID, "bad world / rank (at most 8 peers)");
    CLOVER_REQUIRE(rows_local % 64u == 0 && row0 % 64u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID,
                   "shards are whole 64-row blocks; cols a multiple of 128");
    CLOVER_REQUIRE(rows_local > 0, CLOVER_ERR_UNSUPPORTED, "every rank must own at least one 64-row block (a rank without work could not signal)");
    CLOVER_REQUIRE((reinterpret_cast<uintptr_t>(values_local) & 15u) == 0, CLOVER_ERR_UNSUPPORTED, "values_local must be 16-byte aligned");
    PeerOut peers;
    peers.world = world; peers.rank = rank; peers.epoch = epoch; peers.ticket = ticket; peers.defer = defer;
    for (int p = 0; p < world; ++p) {
        CLOVER_REQUIRE(peer_yv_host[p] && peer_ys_host[p] && peer_flags_host[p], CLOVER_ERR_INVALID, "null peer pointer");
        peers.yv[p] = peer_yv_host[p]; peers.ys[p] = peer_ys_host[p]; peers.flags[p] = peer_flags_host[p];
    }
    // own slice goes through the ordinary local pointers; the stream position of key_host follows clover_m4_mvm_shard
    return launch_mvm<4>(values_local, scales_local, rows_local, cols, row0, xv, xs, nullptr, peers.yv[rank], peers.ys[rank],
                         key_host, (cudaStream_t)stream, &peers);
}

int clover_m4_mvm_shard_fused(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                              uint64_t row0, const int8_t *xv, const float *xs, int8_t *const *peer_yv_host,
                              float *const *peer_ys_host, uint32_t *const *peer_flags_host, unsigned int *ticket,
                              int world, int rank, uint32_t epoch, uint64_t *key_host, void *stream) {
    return launch_shard_fused(values_local, scales_local, rows_local, cols, row0, xv, xs, peer_yv_host, peer_ys_host,
                              peer_flags_host, ticket, world, rank, epoch, key_host, stream, kXchgFlags);
}

int clover_m4_mvm_shard_stamped(const int8_t *values_local, const float *scales_local, uint64_t rows_local, uint64_t cols,
                                uint64_t row0, const int8_t *xv, const float *xs, int8_t *yv_full, float *ys_full,
                                uint64_t *const *peer_msg_host, uint32_t *const *peer_started_host,
                                int world, int rank, uint32_t epoch, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(values_local && scales_local && xv && xs && yv_full && ys_full && peer_msg_host && peer_started_host,
                   CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, CLOVER_ERR_INVALID, "bad world / rank (at most 8 peers)");
    CLOVER_REQUIRE(rows_local % 64u == 0 && row0 % 64u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID,
                   "shards are whole 64-row blocks; cols a multiple of 128");
    CLOVER_REQUIRE(rows_local > 0, CLOVER_ERR_UNSUPPORTED, "every rank must own at least one 64-row block");
    CLOVER_REQUIRE((reinterpret_cast<uintptr_t>(values_local) & 15u) == 0, CLOVER_ERR_UNSUPPORTED, "values_local must be 16-byte aligned");
    PeerOut peers;
    peers.world = world; peers.rank = rank; peers.epoch = epoch; peers.defer = kXchgStamped;
    for (int p = 0; p < world; ++p) {
        CLOVER_REQUIRE(peer_msg_host[p] && peer_started_host[p], CLOVER_ERR_INVALID, "null peer pointer");
        CLOVER_REQUIRE((reinterpret_cast<uintptr_t>(peer_msg_host[p]) & 7u) == 0, CLOVER_ERR_INVALID, "message areas must be 8-byte aligned");
        peers.msg[p] = peer_msg_host[p]; peers.started[p] = peer_started_host[p];
    }
    return launch_mvm<4>(values_local, scales_local, rows_local, cols, row0, xv, xs, nullptr, yv_full, ys_full,
                         key_host, (cudaStream_t)stream, &peers);
}

int clover_m4_shard_stamped_unpack(const uint64_t *msg_local, uint64_t rows, uint64_t row0, uint64_t rows_local, uint32_t epoch,
                                   int8_t *yv_full, float *ys_full, void *stream) {
    CLOVER_REQUIRE(msg_local && yv_full && ys_full, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 64u == 0 && row0 % 64u == 0 && rows_local % 64u == 0 && row0 + rows_local <= rows, CLOVER_ERR_INVALID,
                   "rows, row0 and rows_local are whole 64-row blocks");
    const uint64_t nblocks = rows >> 6;
    if (nblocks == 0 || rows_local == rows) return CLOVER_OK;
    const uint64_t want = (nblocks * 9 + 255) / 256, cap = (uint64_t)sm_count() * 8;
    k_unpack_stamped<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        msg_local, nblocks, row0 >> 6, rows_local >> 6, epoch, reinterpret_cast<uint32_t *>(yv_full), ys_full);
    count_launch();
    return launch_status("k_unpack_stamped");
}

int clover_v4_requantize_mvm(const float *y32, uint64_t rows, int8_t *yv, float *ys, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(y32 && yv && ys, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0, CLOVER_ERR_INVALID, "rows must be a multiple of 128");
    const uint64_t nrb = rows >> 6;
    if (nrb == 0) return CLOVER_OK;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    const unsigned grid = (unsigned)(nrb < cap ? nrb : cap);
    if (key_host) {
        const uint64_t *tables = device_jump_tables();
        if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
        k_requantize_mvm<4, true><<<grid, 64, 0, (cudaStream_t)stream>>>(y32, nrb, yv, ys, key_lanes(key_host), tables);
        host_key_skip(key_host, 2 * nrb);
    } else {
        k_requantize_mvm<4, false><<<grid, 64, 0, (cudaStream_t)stream>>>(y32, nrb, yv, ys, Key4{}, nullptr);
    }
    count_launch();
    return launch_status("k_requantize_mvm");
}

}  // extern "C"
