// vector_kernels.cu - CloverVector4 / CloverVector8: quantize, restore, dot.
//
//   quantize : fp32 -> {4,8}-bit + one fp32 absmax per 64 elements. HBM-bound streaming pass
//              (4.5625 B/elem for 4-bit, 5.0625 B/elem for 8-bit).
//              Reference: include/CloverVector4.h:605-807, include/CloverVector8.h:393-605.
//   restore  : include/CloverVector4.h:1027-1093, include/CloverVector8.h:835-909.
//   dot      : include/CloverVector4.h:1095-1192, include/CloverVector8.h:911-977.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <utility>
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

// =============================================================================================
// quantize
// =============================================================================================
// ---------------------------------------------------------------------------------------------
// quantize: FOUR threads per block of 64 (16 contiguous elements each), no shared memory, no
// CTA barrier. Thread s of a group loads its 64 bytes with four 16 B loads (the two halves of each 32 B sector
// are fetched by consecutive instructions and meet in L1), the block absmax is two xor-shuffles away, and the
// thread's 16 roundings pack into exactly the 8 (4-bit) or 16 (8-bit) contiguous bytes it stores - a warp writes
// 256 / 512 contiguous bytes. ~45 registers => full occupancy; this is what lifted C2a from 73 % of the HBM
// roofline (a thread-per-block kernel, round 1a) - see DESIGN.md section 3.
// Stochastic mode: a group walks R consecutive blocks and all four threads step the same XORShift state
// (two calls per block), so the stream is consumed exactly like the reference's sequential loop.
// ---------------------------------------------------------------------------------------------
constexpr int kQ4Threads = 256;

template <int BITS, bool STOCH>
__global__ void __launch_bounds__(kQ4Threads)
k_vquantize4t(const float *__restrict__ x, uint64_t nblocks, uint64_t R, int8_t *__restrict__ values,
              float *__restrict__ scales, Key4 key, const uint64_t *__restrict__ tables) {
    constexpr float kQmax = BITS == 4 ? 7.0f : 127.0f;
    const int s = threadIdx.x & 3;
    const uint64_t group = ((uint64_t)blockIdx.x * kQ4Threads + threadIdx.x) >> 2;
    const uint64_t ngroups = ((uint64_t)gridDim.x * kQ4Threads) >> 2;
    // rounding disabled: grid-stride (a warp covers 8 consecutive blocks = 2 KiB); stochastic: R consecutive blocks
    uint64_t blk = STOCH ? group * R : group;
    const uint64_t step = STOCH ? 1 : ngroups;
    const uint64_t end = STOCH ? min(nblocks, (group + 1) * R) : nblocks;
    uint64_t lanes[4] = {0, 0, 0, 0};
    if (STOCH && blk < end) {
#pragma unroll
        for (int k = 0; k < 4; ++k) lanes[k] = xs_jump(tables, key.x[k], 2 * blk);
    }
    // whole warps iterate together (shuffles): the trip count is decided per GROUP, groups of a warp may differ
    for (;; blk += step) {
        const bool live = blk < end;
        if (!__any_sync(0xFFFFFFFFu, live)) break;
        float f[16];
        if (live) {
            const float4 *src = reinterpret_cast<const float4 *>(x + blk * 64 + 16 * s);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = __ldg(src + j);
                f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = 0.f;
        }
        float m = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) m = fmaxf(m, fabsf(f[e]));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
        m = guard_zero(m);
        const float scale = quant_scale(kQmax, m);
        uint32_t w[8];                                   // the PRNG call this thread's elements use (call s / 2)
        if (STOCH) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t o = xs_next(lanes[k]);
                    if (c == (s >> 1)) { w[2 * k] = (uint32_t)o; w[2 * k + 1] = (uint32_t)(o >> 32); }
                }
        }
        if (!live) continue;
        if (s == 0) scales[blk] = m;
        int q[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            // element e = 16 s + i: noise slot (call e / 32, byte (e % 32) / 8, word e % 8)
            const float rnd = STOCH ? noise_from_word(w[i & 7], 2 * (s & 1) + (i >> 3)) : 0.f;
            q[i] = quant_one(f[i], scale, rnd);
        }
        if (BITS == 4) {
            *reinterpret_cast<uint2 *>(values + blk * 32 + 8 * s) = make_uint2(pack8_nibbles(q), pack8_nibbles(q + 8));
        } else {
            *reinterpret_cast<uint4 *>(values + blk * 64 + 16 * s) =
                make_uint4(pack4_bytes(q), pack4_bytes(q + 4), pack4_bytes(q + 8), pack4_bytes(q + 12));
        }
    }
}

template <int BITS>
static int launch_vquantize(const float *x, uint64_t n_pad, int8_t *values, float *scales, uint64_t *key_host,
                            cudaStream_t stream) {
    const uint64_t nblocks = n_pad / kBlock;
    if (nblocks == 0) return CLOVER_OK;
    Key4 key = {};
    const uint64_t *tables = nullptr;
    if (key_host) {
        tables = device_jump_tables();
        if (!tables) { set_error("clover: could not upload PRNG jump tables"); return CLOVER_ERR_CUDA; }
        key = key_lanes(key_host);
    }
    {
        static int ctas_per_sm[kMaxDevices][2] = {};
        const int dev = current_device() < 0 ? 0 : current_device();
        int &cps = ctas_per_sm[dev][key_host != nullptr];
        if (cps == 0) {
            cudaError_t e = key_host ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, k_vquantize4t<BITS, true>, kQ4Threads, 0)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, k_vquantize4t<BITS, false>, kQ4Threads, 0);
            if (e != cudaSuccess || cps < 1) cps = 1;
        }
        const uint64_t max_groups = (uint64_t)sm_count() * cps * (kQ4Threads / 4);       // one resident wave
        const uint64_t groups = nblocks < max_groups ? nblocks : max_groups;
        const uint64_t R = (nblocks + groups - 1) / groups;
        const unsigned grid = (unsigned)((groups * 4 + kQ4Threads - 1) / kQ4Threads);
        if (key_host) k_vquantize4t<BITS, true><<<grid, kQ4Threads, 0, stream>>>(x, nblocks, R, values, scales, key, tables);
        else          k_vquantize4t<BITS, false><<<grid, kQ4Threads, 0, stream>>>(x, nblocks, R, values, scales, key, nullptr);
    }
    if (key_host) host_key_skip(key_host, 2 * nblocks);
    count_launch();
    return launch_status("k_vquantize");
}

// =============================================================================================
// restore
// =============================================================================================
__global__ void __launch_bounds__(256)
k_v4_restore(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t nwords, float *__restrict__ x) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        const float s = __fdiv_rn(scales[i >> 3], 7.0f);        // su[b] / 7.0f (:1050)
        const uint32_t w = values[i];
        float o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = (int)(int8_t)(w >> (8 * j));
            o[2 * j] = __fmul_rn(__int2float_rn(b >> 4), s);
            o[2 * j + 1] = __fmul_rn(__int2float_rn((int)(int8_t)(b << 4) >> 4), s);
        }
        float4 *dst = reinterpret_cast<float4 *>(x + i * 8);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

__global__ void __launch_bounds__(256)
k_v8_restore(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t nwords, float *__restrict__ x) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        const float s = __fdiv_rn(scales[i >> 4], 127.0f);      // su[b] / 127.0f (:872)
        const uint32_t w = values[i];
        float4 o;
        o.x = __fmul_rn(__int2float_rn((int)(int8_t)(w)), s);
        o.y = __fmul_rn(__int2float_rn((int)(int8_t)(w >> 8)), s);
        o.z = __fmul_rn(__int2float_rn((int)(int8_t)(w >> 16)), s);
        o.w = __fmul_rn(__int2float_rn((int)(int8_t)(w >> 24)), s);
        *reinterpret_cast<float4 *>(x + i * 4) = o;
    }
}

// =============================================================================================
// dot - exact-order mode
// =============================================================================================
// One warp. 4-bit: lane = 8*a + l for the reference's accumulator a (even/odd blocks) and AVX lane
// l; each of the 16 lanes runs its fp32 FMA chain in block order, then acc1+acc2 and the hadd tree.
// 8-bit: 8 lanes, one accumulator (CloverVector8.h:911-977).
template <int BITS>
__global__ void __launch_bounds__(32)
k_vdot_exact(const uint32_t *__restrict__ u, const float *__restrict__ su, const uint32_t *__restrict__ v,
             const float *__restrict__ sv, uint64_t nblocks, float *__restrict__ result) {
    const int lane = threadIdx.x;
    float acc = 0.f;
    if (BITS == 4) {
        const int a = (lane >> 3) & 1, l = lane & 7;
        if (lane < 16) {
            const float rcp49 = 1.0f / 49.0f;
#pragma unroll 4
            for (uint64_t b = a; b < nblocks; b += 2) {
                const int dot = nibble_dot_word(u[b * 8 + l], v[b * 8 + l]);
                const float s = __fmul_rn(__fmul_rn(su[b], rcp49), sv[b]);
                acc = __fmaf_rn(s, __int2float_rn(dot), acc);
            }
        }
        acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, 8));   // acc_1 + acc_2 (:1190)
    } else {
        const int l = lane & 7;
        if (lane < 8) {
            const float rcp127 = 1.0f / 127.0f;
#pragma unroll 4
            for (uint64_t b = 0; b < nblocks; ++b) {
                int dot = dp4a_ss((int)u[b * 16 + l], (int)v[b * 16 + l], 0);
                dot = dp4a_ss((int)u[b * 16 + 8 + l], (int)v[b * 16 + 8 + l], dot);
                const float s = __fmul_rn(__fmul_rn(su[b], rcp127), __fmul_rn(sv[b], rcp127));
                acc = __fmaf_rn(s, __int2float_rn(dot), acc);
            }
        }
    }
    acc = hadd8_butterfly(acc);
    if (lane == 0) *result = acc;
}

// The same chains for up to 1024 blocks (the AUTO range), latency-optimised: the per-(block, lane) integers and the per-block
// scale products do not depend on the summation order, so a whole CTA computes them in parallel with coalesced loads (one
// or a few memory round trips) into shared memory; warp 0 then runs the reference's fp32 chains out of shared memory. The
// one-warp kernel above pays a dependent global round trip every four blocks: 9.6 us at n = 4096, ~90 us at n = 65536.
constexpr int kDotExactBlocks = 1024;
template <int BITS>
__global__ void __launch_bounds__(256)
k_vdot_exact_cta(const uint32_t *__restrict__ u, const float *__restrict__ su, const uint32_t *__restrict__ v,
                 const float *__restrict__ sv, uint32_t nblocks, float *__restrict__ result) {
    __shared__ int dots[kDotExactBlocks * 8];
    __shared__ float prod[kDotExactBlocks];
    const int tid = threadIdx.x;
    for (uint32_t idx = tid; idx < nblocks * 8; idx += 256) {
        if (BITS == 4) {
            dots[idx] = nibble_dot_word(__ldg(u + idx), __ldg(v + idx));             // word (b, l) = u[b * 8 + l]
        } else {
            const uint32_t b = idx >> 3, l = idx & 7;
            int d = dp4a_ss((int)__ldg(u + b * 16 + l), (int)__ldg(v + b * 16 + l), 0);
            dots[idx] = dp4a_ss((int)__ldg(u + b * 16 + 8 + l), (int)__ldg(v + b * 16 + 8 + l), d);
        }
    }
    for (uint32_t b = tid; b < nblocks; b += 256) {
        if (BITS == 4) prod[b] = __fmul_rn(__fmul_rn(__ldg(su + b), 1.0f / 49.0f), __ldg(sv + b));
        else           prod[b] = __fmul_rn(__fmul_rn(__ldg(su + b), 1.0f / 127.0f), __fmul_rn(__ldg(sv + b), 1.0f / 127.0f));
    }
    __syncthreads();
    if (tid >= 32) return;
    const int lane = tid, l = lane & 7;
    float acc = 0.f;
    if (BITS == 4) {
        const int a = (lane >> 3) & 1;
        if (lane < 16) {
#pragma unroll 8
            for (uint32_t b = a; b < nblocks; b += 2) acc = __fmaf_rn(prod[b], __int2float_rn(dots[b * 8 + l]), acc);
        }
        acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, 8));   // acc_1 + acc_2 (:1190)
    } else {
        if (lane < 8) {
#pragma unroll 8
            for (uint32_t b = 0; b < nblocks; ++b) acc = __fmaf_rn(prod[b], __int2float_rn(dots[b * 8 + l]), acc);
        }
    }
    acc = hadd8_butterfly(acc);
    if (lane == 0) *result = acc;
}

// =============================================================================================
// dot - fast mode
// =============================================================================================
// Grid-stride over 16-byte chunks of both operands (HBM/L2-bound: 1.125 B/elem for 4-bit). The
// per-block integer is exact; each block contributes (double)s_b * I_b with the reference's fp32
// scale s_b = (su*(1/49))*sv, summed in fp64 by a fixed tree (thread -> warp -> CTA -> last CTA),
// so the result is deterministic and equals the fp64 evaluation of the reference's own terms.
constexpr int kDotThreads = 256;

struct DotWorkspace { double *partials; unsigned int *ticket; int capacity; };

template <int BITS, int UNROLL>
__global__ void __launch_bounds__(kDotThreads)
k_vdot_fast(const uint4 *__restrict__ u, const float *__restrict__ su, const uint4 *__restrict__ v,
            const float *__restrict__ sv, uint64_t nchunks, double *__restrict__ partials,
            unsigned int *__restrict__ ticket, float *__restrict__ result) {
    constexpr int kChunksPerBlock = BITS == 4 ? 2 : 4;      // 16 B chunks per 64-element block
    const uint64_t stride = (uint64_t)gridDim.x * kDotThreads;
    double acc = 0.0;
    // all 32 lanes of a warp run the same trip count (nchunks is a multiple of 32: n_pad % 128 == 0
    // gives nchunks % 4 == 0 only, so guard lanes individually but keep shuffles warp-uniform).
    // UNROLL chunks per operand and thread are requested before the first is consumed.
    const uint64_t first = (uint64_t)blockIdx.x * kDotThreads + threadIdx.x;
    const uint64_t warp_first = first - (threadIdx.x & 31);
    // launched with programmatic stream serialization: the grid may be set up while its predecessor in the stream drains;
    // nothing of the predecessor's output (operands, the ticket, the result word) is touched before this returns
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (uint64_t base = warp_first; base < nchunks; base += stride * UNROLL) {
        uint4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const uint64_t i = base + (uint64_t)k * stride + (threadIdx.x & 31);
            if (i < nchunks) { a[k] = ldg_stream(u + i); b[k] = ldg_stream(v + i); }
            else             { a[k] = make_uint4(0u, 0u, 0u, 0u); b[k] = a[k]; }
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const uint64_t i = base + (uint64_t)k * stride + (threadIdx.x & 31);
            int part;
            if (BITS == 4) {
                part = nibble_dot_word(a[k].x, b[k].x) + nibble_dot_word(a[k].y, b[k].y) + nibble_dot_word(a[k].z, b[k].z) +
                       nibble_dot_word(a[k].w, b[k].w);
            } else {
                part = dp4a_ss((int)a[k].x, (int)b[k].x, 0);
                part = dp4a_ss((int)a[k].y, (int)b[k].y, part);
                part = dp4a_ss((int)a[k].z, (int)b[k].z, part);
                part = dp4a_ss((int)a[k].w, (int)b[k].w, part);
            }
            part += __shfl_xor_sync(0xFFFFFFFFu, part, 1);
            if (BITS == 8) part += __shfl_xor_sync(0xFFFFFFFFu, part, 2);
            if (i < nchunks && (i % kChunksPerBlock) == 0) {
                const uint64_t blk = i / kChunksPerBlock;
                float s;
                if (BITS == 4) s = __fmul_rn(__fmul_rn(su[blk], 1.0f / 49.0f), sv[blk]);
                else           s = __fmul_rn(__fmul_rn(su[blk], 1.0f / 127.0f), __fmul_rn(sv[blk], 1.0f / 127.0f));
                acc += (double)s * (double)part;
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // a dependent grid may be scheduled from here on (it waits at its top)
    // fixed reduction tree
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __shared__ double warp_sums[kDotThreads / 32];
    __shared__ bool is_last;
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kDotThreads / 32; ++w) s += warp_sums[w];
        partials[blockIdx.x] = s;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += kDotThreads) s += __ldcg(partials + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kDotThreads / 32; ++w) t += warp_sums[w];
            *result = (float)t;
            *ticket = 0u;                                   // self-cleaning for the next call
        }
    }
}

// per-(device, stream) workspace for the cross-CTA reduction
static int dot_workspace(cudaStream_t stream, int grid, DotWorkspace *out) {
    void *p = nullptr, *t = nullptr;
    int rc = stream_scratch(kScratchDotPartials, stream, sizeof(double) * (size_t)grid, 0, &p);
    if (rc != CLOVER_OK) return rc;
    rc = stream_scratch(kScratchDotTicket, stream, sizeof(unsigned int), sizeof(unsigned int), &t);
    if (rc != CLOVER_OK) return rc;
    out->partials = static_cast<double *>(p);
    out->ticket = static_cast<unsigned int *>(t);
    out->capacity = grid;
    return CLOVER_OK;
}

constexpr uint64_t kDotExactLimit = 1ull << 16;   // AUTO: exact-order chains up to 65536 elements

template <int BITS>
static int launch_vdot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n_pad,
                       float *result, int mode, cudaStream_t stream) {
    const uint64_t nblocks = n_pad / kBlock;
    if (mode == CLOVER_DOT_AUTO) mode = (n_pad <= kDotExactLimit) ? CLOVER_DOT_EXACT : CLOVER_DOT_FAST;
    if (mode == CLOVER_DOT_EXACT || nblocks == 0) {
        if (nblocks > 0 && nblocks <= (uint64_t)kDotExactBlocks)
            k_vdot_exact_cta<BITS><<<1, 256, 0, stream>>>(reinterpret_cast<const uint32_t *>(u), su,
                                                          reinterpret_cast<const uint32_t *>(v), sv, (uint32_t)nblocks, result);
        else
        k_vdot_exact<BITS><<<1, 32, 0, stream>>>(reinterpret_cast<const uint32_t *>(u), su,
                                                 reinterpret_cast<const uint32_t *>(v), sv, nblocks, result);
        count_launch();
        return launch_status("k_vdot_exact");
    }
    const uint64_t nchunks = BITS == 4 ? n_pad / 32 : n_pad / 16;
    uint64_t want = (nchunks + kDotThreads * 4 - 1) / (kDotThreads * 4);       // >= 4 chunks per thread
    // 8 CTAs per SM, two chunks per operand and thread in flight: measured on B200 at n = 2^26 (tools/dot_sweep.py, us for
    // unroll 1 / 2 / 4): 4-bit 16.4 / 16.4 / 18.5, 8-bit 24.3 / 23.3 / 26.7; 2 or 4 CTAs per SM are slower for every unroll
    const uint64_t cap = (uint64_t)sm_count() * 8;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    DotWorkspace ws;
    int rc = dot_workspace(stream, (int)cap, &ws);
    if (rc != CLOVER_OK) return rc;
    // programmatic dependent launch: of the 16.4 us a 2^26-element 4-bit dot takes, ~5 are launch ramp and the last-CTA
    // finish; back-to-back dots overlap the next grid's set-up with the finish of the previous one
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kDotThreads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CLOVER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_vdot_fast<BITS, 2>, reinterpret_cast<const uint4 *>(u), su,
                                         reinterpret_cast<const uint4 *>(v), sv, nchunks, ws.partials, ws.ticket, result));
    count_launch();
    return launch_status("k_vdot_fast");
}

}  // namespace clover

using namespace clover;

#define CLOVER_CHECK_VEC(n_pad)                                                                     \
    CLOVER_REQUIRE((n_pad) % 128u == 0, CLOVER_ERR_INVALID, "n_pad must be a multiple of 128 (clover_size_pad)")

extern "C" {

uint64_t clover_dot_exact_limit(void) { return kDotExactLimit; }

int clover_v4_quantize(const float *x, uint64_t n_pad, int8_t *values, float *scales, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(x && values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_VEC(n_pad);
    return launch_vquantize<4>(x, n_pad, values, scales, key_host, (cudaStream_t)stream);
}
int clover_v8_quantize(const float *x, uint64_t n_pad, int8_t *values, float *scales, uint64_t *key_host, void *stream) {
    CLOVER_REQUIRE(x && values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_VEC(n_pad);
    return launch_vquantize<8>(x, n_pad, values, scales, key_host, (cudaStream_t)stream);
}

int clover_v4_restore(const int8_t *values, const float *scales, uint64_t n_pad, float *x, void *stream) {
    CLOVER_REQUIRE(x && values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_VEC(n_pad);
    const uint64_t nwords = n_pad / 8;
    if (nwords == 0) return CLOVER_OK;
    const uint64_t want = (nwords + 255) / 256, cap = (uint64_t)sm_count() * 16;
    k_v4_restore<<<(unsigned)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t *>(values), scales, nwords, x);
    count_launch();
    return launch_status("k_v4_restore");
}
int clover_v8_restore(const int8_t *values, const float *scales, uint64_t n_pad, float *x, void *stream) {
    CLOVER_REQUIRE(x && values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_CHECK_VEC(n_pad);
    const uint64_t nwords = n_pad / 4;
    if (nwords == 0) return CLOVER_OK;
    const uint64_t want = (nwords + 255) / 256, cap = (uint64_t)sm_count() * 16;
    k_v8_restore<<<(unsigned)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t *>(values), scales, nwords, x);
    count_launch();
    return launch_status("k_v8_restore");
}

int clover_v4_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n_pad,
                  float *result, int mode, void *stream) {
    CLOVER_REQUIRE(u && su && v && sv && result, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(mode >= 0 && mode <= 2, CLOVER_ERR_INVALID, "bad dot mode");
    CLOVER_CHECK_VEC(n_pad);
    return launch_vdot<4>(u, su, v, sv, n_pad, result, mode, (cudaStream_t)stream);
}
int clover_v8_dot(const int8_t *u, const float *su, const int8_t *v, const float *sv, uint64_t n_pad,
                  float *result, int mode, void *stream) {
    CLOVER_REQUIRE(u && su && v && sv && result, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(mode >= 0 && mode <= 2, CLOVER_ERR_INVALID, "bad dot mode");
    CLOVER_CHECK_VEC(n_pad);
    return launch_vdot<8>(u, su, v, sv, n_pad, result, mode, (cudaStream_t)stream);
}

}  // extern "C"
