// threshold_kernels.cu - CloverVector4/8::threshold(k): keep the k elements of largest magnitude, zero the rest, in place
// (include/CloverVector4.h:1913-1973, include/CloverVector8.h:1680-1740).
//
// The magnitude of element i is the reference's getAbs(i), reproduced bit for bit:
//   4-bit: | (scale / 7.0f) * float(q) |          (include/CloverVector4.h:190-203)
//   8-bit: | (float(q) * scale) / 127.0f |        (include/CloverVector8.h:141-147)
// Magnitudes are non-negative fp32, so their bit patterns order like unsigned integers.
//
// Two paths (clover_threshold_mode):
//   EXACT  the reference's own sequential walk - std::make_heap over the first k elements (libstdc++ __make_heap /
//          __adjust_heap / __push_heap), then "replace the root if STRICTLY larger" + min_heapify
//          (include/CloverBase.h:226-249) - executed by one thread over magnitudes a parallel kernel prepared. Which of
//          several equal magnitudes survives depends on the heap layout; this path leaves exactly the reference's bytes.
//          Latency-bound, meant for parity runs (n <= clover_threshold_exact_limit()).
//   FAST   an 8-bit-digit radix select over the magnitude bits (4 histogram passes, all on the device, no host sync; one
//          CTA up to 8192 elements, one 8-CTA cluster merging its histograms through DSMEM up to 262144, 7 launches beyond)
//          finds the k-th largest magnitude t; everything above t stays, everything below goes, and of the elements equal
//          to t the ones with the LOWEST indices stay until k survivors are reached (the reference's sequential and
//          OpenMP variants already disagree with each other on such ties; its acceptance test - 02_vector.cpp:450-500 -
//          compares sorted magnitudes only, which this path satisfies exactly).
#include <stdlib.h>
#include <algorithm>
#include <mutex>
#include <cooperative_groups.h>
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

constexpr int kThrThreads = 256;
constexpr int kThrSmallThreads = 1024;
constexpr uint64_t kThrExactLimit = 1u << 12;     // AUTO: the sequential heap walk up to 4096 elements
constexpr uint64_t kThrSmallLimit = 1u << 13;     // FAST: one CTA does the whole selection up to 8192 elements,
constexpr uint64_t kThrClusterLimit = 1u << 18;   //       one 8-CTA cluster (histograms merged through DSMEM) up to 262144 (14 us at 32768, 22-27 us at 131072; at 2^20 the seven-launch path wins 40 : 114 us)
constexpr int kThrClusterSize = 8;

struct ThrState {            // device-resident selection state; hist[] and ticket are zero between calls
    uint32_t prefix;         // magnitude bits decided so far (high digits)
    uint32_t mask;           // which bits of `prefix` are decided
    uint64_t k_rem;          // survivors still to be found among the elements matching prefix
    uint32_t ticket;         // CTAs that have finished the current pass
    uint32_t hist[256];
};

// Histogram update. Measured on B200 at n = 2^26 (random and all-equal inputs alike): merging equal bins across the warp
// first (match.any / match.all) is 4-17 % SLOWER than plain shared-memory atomics - the passes are bound by the ~30
// instructions per element that rebuild the magnitude, not by bin contention - so the atomics stay plain.
__device__ __forceinline__ void hist_add(uint32_t *h, uint32_t digit, bool valid) {
    if (valid) atomicAdd(&h[digit], 1u);
}

// the digit in which the k-th largest of the counted elements falls: returns it, `above` = elements with a larger digit.
// Called by every thread of a CTA with >= 256 threads; h[] is shared memory holding the full histogram.
__device__ __forceinline__ uint32_t pick_digit(const uint32_t *h, uint64_t k, uint64_t *scratch /* 256 */, uint64_t &above) {
    const uint32_t d = threadIdx.x;
    if (d < 256) scratch[d] = h[d];
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {                      // inclusive suffix sums: scratch[d] = sum_{j >= d} h[j]
        uint64_t v = 0;
        if (d < 256 && d + o < 256) v = scratch[d + o];
        __syncthreads();
        if (d < 256) scratch[d] += v;
        __syncthreads();
    }
    __shared__ uint32_t picked;
    __shared__ uint64_t picked_above;
    if (d < 256) {
        const uint64_t ge = scratch[d], gt = ge - h[d];
        if (gt < k && k <= ge) { picked = d; picked_above = gt; }        // exactly one digit satisfies this
    }
    __syncthreads();
    above = picked_above;
    return picked;
}

// magnitudes of the 8 elements of a 4-bit word (element order) / the 4 elements of an 8-bit word, as ordered bit patterns
__device__ __forceinline__ uint32_t abs_bits4(float scale7, int q) { return __float_as_uint(fabsf(__fmul_rn(scale7, __int2float_rn(q)))); }
__device__ __forceinline__ uint32_t abs_bits8(float scale, int q) { return __float_as_uint(fabsf(__fdiv_rn(__fmul_rn(__int2float_rn(q), scale), 127.0f))); }

// element e (0..7) of a 4-bit word: byte e>>1, even elements in the HIGH nibble
__device__ __forceinline__ int nib_of(uint32_t w, int e) {
    const int sh = 8 * (e >> 1) + ((e & 1) ? 0 : 4);
    return ((int)(w << (28 - sh))) >> 28;
}

template <int BITS> struct ThrWord {
    static constexpr int kElems = BITS == 4 ? 8 : 4;     // elements per 32-bit word
    __device__ static void mags(uint32_t w, float scale, uint64_t first, uint64_t n, uint32_t *m) {
        const float s = BITS == 4 ? __fdiv_rn(scale, 7.0f) : scale;
#pragma unroll
        for (int e = 0; e < kElems; ++e) {
            const int q = BITS == 4 ? nib_of(w, e) : (int)(int8_t)(w >> (8 * e));
            const uint32_t b = BITS == 4 ? abs_bits4(s, q) : abs_bits8(s, q);
            m[e] = first + e < n ? b : 0u;                // pad elements never compete
        }
    }
    __device__ static uint32_t clear(uint32_t w, int e) {
        return BITS == 4 ? w & ~(0xFu << (8 * (e >> 1) + ((e & 1) ? 0 : 4))) : w & ~(0xFFu << (8 * e));
    }
};

// ---- FAST (large n): one histogram pass over digit `shift` of the magnitudes that match the decided prefix; the last
// CTA to finish extends the prefix by the digit in which the k_rem-th largest matching element falls -------------------
template <int BITS, bool TOP>
__global__ void __launch_bounds__(kThrThreads)
k_thr_hist(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint64_t nwords, int shift,
           uint64_t k, ThrState *__restrict__ st) {
    constexpr int E = ThrWord<BITS>::kElems;
    __shared__ uint32_t h[256];
    __shared__ uint64_t scratch[256];
    __shared__ bool is_last;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = TOP ? 0u : st->prefix, mask = TOP ? 0u : st->mask;
    const uint64_t stride = (uint64_t)gridDim.x * kThrThreads;
    const uint64_t warp_first = (uint64_t)blockIdx.x * kThrThreads + (threadIdx.x & ~31u);
    for (uint64_t base = warp_first; base < nwords; base += stride) {         // warp-uniform trip count (collectives inside)
        const uint64_t i = base + (threadIdx.x & 31);
        uint32_t m[E];
        if (i < nwords) ThrWord<BITS>::mags(values[i], scales[(i * E) >> 6], i * E, n, m);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool valid = i < nwords && i * E + e < n && (m[e] & mask) == prefix;
            hist_add(h, valid ? (m[e] >> shift) & 0xFFu : 0u, valid);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    h[threadIdx.x] = __ldcg(&st->hist[threadIdx.x]);
    __syncthreads();
    const uint64_t k_rem = TOP ? k : st->k_rem;
    uint64_t above;
    const uint32_t d = pick_digit(h, k_rem, scratch, above);
    st->hist[threadIdx.x] = 0;                                                // leave the state clean for the next pass / call
    if (threadIdx.x == 0) {
        st->prefix = prefix | (d << shift);
        st->mask = mask | (0xFFu << shift);
        st->k_rem = k_rem - above;
        st->ticket = 0;
    }
}

// ---- FAST: ties at the threshold. Each CTA owns one contiguous range of words (index order) ------------------------
template <int BITS>
__global__ void __launch_bounds__(kThrThreads)
k_thr_count_ties(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint64_t nwords,
                 uint64_t words_per_cta, const ThrState *__restrict__ st, uint32_t *__restrict__ tie_count) {
    constexpr int E = ThrWord<BITS>::kElems;
    const uint32_t t = st->prefix;
    const uint64_t w0 = (uint64_t)blockIdx.x * words_per_cta, w1 = min(w0 + words_per_cta, nwords);
    uint32_t c = 0;
    for (uint64_t i = w0 + threadIdx.x; i < w1; i += kThrThreads) {
        uint32_t m[E];
        ThrWord<BITS>::mags(values[i], scales[(i * E) >> 6], i * E, n, m);
#pragma unroll
        for (int e = 0; e < E; ++e) c += (i * E + e < n && m[e] == t) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    __shared__ uint32_t ws[kThrThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kThrThreads / 32; ++w) s += ws[w];
        tie_count[blockIdx.x] = s;
    }
}

// one block: exclusive prefix sums of the per-CTA tie counts (64-bit: n may exceed 2^32)
__global__ void __launch_bounds__(256) k_thr_scan(const uint32_t *__restrict__ tie_count, uint64_t *__restrict__ tie_base, int nctas) {
    __shared__ uint64_t part[256];
    const int per = (nctas + 255) / 256, a = threadIdx.x * per, b = min(a + per, nctas);
    uint64_t s = 0;
    for (int i = a; i < b; ++i) s += tie_count[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int i = 0; i < 256; ++i) { const uint64_t v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (int i = a; i < b; ++i) { tie_base[i] = run; run += tie_count[i]; }
}

template <int E> __device__ __forceinline__ void load_mags(const uint32_t *__restrict__ mag, uint64_t word, uint32_t *m) {
#pragma unroll
    for (int j = 0; j < E / 4; ++j) {
        const uint4 v = reinterpret_cast<const uint4 *>(mag)[word * (E / 4) + j];
        m[4 * j] = v.x; m[4 * j + 1] = v.y; m[4 * j + 2] = v.z; m[4 * j + 3] = v.w;
    }
}

// Clear what does not survive in words [w0, w1), walked in index order by the whole CTA (NT threads): magnitudes above t
// stay, below t go, and an element equal to t stays while fewer than keep_ties such elements precede it (`first_rank` =
// ties before w0).
template <int BITS, int NT>
__device__ __forceinline__ void apply_range(uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint64_t w0,
                                            uint64_t w1, uint32_t t, uint64_t keep_ties, uint64_t first_rank,
                                            const uint32_t *__restrict__ mag = nullptr /* cached magnitudes, E per word */) {
    constexpr int E = ThrWord<BITS>::kElems;
    __shared__ uint32_t ws[NT / 32];
    __shared__ uint64_t running;
    if (threadIdx.x == 0) running = first_rank;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t base = w0; base < w1; base += NT) {
        const uint64_t i = base + threadIdx.x;
        uint32_t w = 0, m[E], ties = 0;
        if (i < w1) {
            w = values[i];
            if (mag) load_mags<E>(mag, i, m);
            else ThrWord<BITS>::mags(w, scales[(i * E) >> 6], i * E, n, m);
#pragma unroll
            for (int e = 0; e < E; ++e) ties += (i * E + e < n && m[e] == t) ? 1u : 0u;
        }
        // exclusive scan of `ties` over the CTA in thread (= index) order
        uint32_t incl = ties;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int x = 0; x < NT / 32; ++x) { const uint32_t v = ws[x]; if (x < warp) before += v; total += v; }
        uint64_t rank = running + before + (incl - ties);
        if (i < w1) {
            uint32_t out = w;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (i * E + e >= n) continue;
                bool keep = m[e] > t;
                if (m[e] == t) { keep = rank < keep_ties; ++rank; }
                if (!keep) out = ThrWord<BITS>::clear(out, e);
            }
            if (out != w) values[i] = out;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
}

template <int BITS>
__global__ void __launch_bounds__(kThrThreads)
k_thr_apply(uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint64_t nwords, uint64_t words_per_cta,
            const ThrState *__restrict__ st, const uint64_t *__restrict__ tie_base) {
    const uint64_t w0 = (uint64_t)blockIdx.x * words_per_cta, w1 = min(w0 + words_per_cta, nwords);
    apply_range<BITS, kThrThreads>(values, scales, n, w0, w1, st->prefix, st->k_rem, tie_base[blockIdx.x]);
}

// ---- FAST, 4-bit, large n: selection over LEVELS instead of elements -------------------------------------------------
// All 64 elements of a block share one scale, so a block holds at most nine distinct magnitudes |scale / 7 * j|,
// j = |q| = 0..8. One pass over the nibbles counts the elements of each level per block (eight byte counters packed in a
// uint64; level 0 = the valid elements that are left); the four radix passes and the tie count then walk 12 bytes per
// BLOCK (counts + scale: 1/3 of the vector, no per-element magnitude rebuild) and add a level's count to the histogram
// with one atomic; the apply pass compares each nibble's level with the threshold. Same selection rule as the
// element-wise path (everything above the k-th largest magnitude stays, ties are kept in index order): 476 -> 143 us at
// n = 2^26 (levels 36, four radix passes 64, ties 14, apply ~25 us). Used beyond 2^18 elements only: a single-CTA version
// of the same scheme measured 14 / 35 us at n = 32768 / 131072 against 13.6 / 24 us of the cluster kernel (r02n).
__device__ __forceinline__ uint32_t nibble_abs8(uint32_t w) {          // |q| of the 8 two's-complement nibbles, SIMD within the word (0..8, no carries)
    const uint32_t sign = (w >> 3) & 0x11111111u;
    return ((w ^ (sign * 0xFu)) + sign);
}
__device__ __forceinline__ uint32_t level_bits(float s7, int j) { return __float_as_uint(fabsf(__fmul_rn(s7, (float)j))); }   // = abs_bits4 for |q| = j

// counts of levels 1..8 of block b in byte lanes 0..7; elements at index >= n do not exist
__device__ __forceinline__ uint64_t block_levels(const uint4 *__restrict__ values, uint64_t n, uint64_t b) {
    const uint4 v0 = values[2 * b], v1 = values[2 * b + 1];
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const uint64_t valid = n - b * 64 < 64 ? n - b * 64 : 64;            // elements of this block below n
    uint64_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t a = nibble_abs8(w[i]);
        // element 8i + e sits in nibble (e ^ 1) of the word (even elements in the HIGH nibble of each byte)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t lv = (a >> (4 * (e ^ 1))) & 0xFu;
            if (lv != 0 && (uint64_t)(8 * i + e) < valid) cnt += 1ull << (8 * (lv - 1));
        }
    }
    return cnt;
}

// thread = block
__global__ void __launch_bounds__(kThrThreads)
k_thr4_levels(const uint4 *__restrict__ values, uint64_t n, uint64_t nblocks, uint64_t *__restrict__ levels) {
    for (uint64_t b = (uint64_t)blockIdx.x * kThrThreads + threadIdx.x; b < nblocks; b += (uint64_t)gridDim.x * kThrThreads)
        levels[b] = block_levels(values, n, b);
}

// element count of level j (0..8) of a block
__device__ __forceinline__ uint32_t level_count(uint64_t cnt, uint32_t valid, int j) {
    if (j) return (uint32_t)(cnt >> (8 * (j - 1))) & 0xFFu;
    uint64_t t = cnt;                                                       // level 0: what is left of the valid elements
    t = (t & 0x00FF00FF00FF00FFull) + ((t >> 8) & 0x00FF00FF00FF00FFull);
    t = (t & 0x0000FFFF0000FFFFull) + ((t >> 16) & 0x0000FFFF0000FFFFull);
    return valid - (uint32_t)((t + (t >> 32)) & 0xFFFFu);
}

// Histogram of digit `shift` of the level magnitudes that match the decided prefix, blocks first, first + stride, ... < nblocks
// (`first` is the position of the calling WARP's lane 0; all 32 lanes call with the same trip count).
// The magnitudes of neighbouring blocks share their high digits, so plain shared-memory atomics would serialise on a
// handful of bins (9 same-address atomics per thread: 130 us of the first version's 184 us at n = 2^26). A thread
// therefore merges the consecutive levels that fall into one bin (levels are monotone in j) and the warp merges equal
// bins across its lanes (match.any + redux) before ONE atomic per distinct bin.
__device__ __forceinline__ void thr4_hist_range(const uint64_t *__restrict__ levels, const float *__restrict__ scales, uint64_t n,
                                                uint64_t nblocks, uint64_t warp_first, uint64_t stride, int shift, uint32_t prefix,
                                                uint32_t mask, uint32_t *h) {
    auto flush = [&](uint32_t digit, uint32_t c) {
        const uint32_t key = c ? digit : 0x100u;                             // 0x100: this lane has nothing to add
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t total = __reduce_add_sync(peers, c);
        if (c && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[digit], total);
    };
    for (uint64_t base = warp_first; base < nblocks; base += stride) {
        const uint64_t b = base + (threadIdx.x & 31);
        const bool live = b < nblocks;
        const uint64_t cnt = live ? levels[b] : 0ull;
        const uint32_t valid = live ? (uint32_t)(n - b * 64 < 64 ? n - b * 64 : 64) : 0u;
        const float s7 = live ? __fdiv_rn(scales[b], 7.0f) : 0.f;
        uint32_t run_digit = 0x100u, run_count = 0;
#pragma unroll
        for (int j = 0; j <= 8; ++j) {
            const uint32_t m = level_bits(s7, j);
            const uint32_t c = (m & mask) == prefix ? level_count(cnt, valid, j) : 0u;
            const uint32_t d = (m >> shift) & 0xFFu;
            // warp-uniform decision: somebody has to start a new run -> everybody flushes (a lane without a change flushes 0)
            const bool change = c && d != run_digit;
            if (__any_sync(0xFFFFFFFFu, change && run_count)) {
                flush(run_digit, change ? run_count : 0u);
                if (change) run_count = 0;
            }
            if (c) { run_digit = d; run_count += c; }
        }
        flush(run_digit, run_count);
    }
}

template <bool TOP>
__global__ void __launch_bounds__(kThrThreads)
k_thr4_hist(const uint64_t *__restrict__ levels, const float *__restrict__ scales, uint64_t n, uint64_t nblocks, int shift,
            uint64_t k, ThrState *__restrict__ st) {
    __shared__ uint32_t h[256];
    __shared__ uint64_t scratch[256];
    __shared__ bool is_last;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = TOP ? 0u : st->prefix, mask = TOP ? 0u : st->mask;
    thr4_hist_range(levels, scales, n, nblocks, (uint64_t)blockIdx.x * kThrThreads + (threadIdx.x & ~31u), (uint64_t)gridDim.x * kThrThreads,
                    shift, prefix, mask, h);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    h[threadIdx.x] = __ldcg(&st->hist[threadIdx.x]);
    __syncthreads();
    const uint64_t k_rem = TOP ? k : st->k_rem;
    uint64_t above;
    const uint32_t d = pick_digit(h, k_rem, scratch, above);
    st->hist[threadIdx.x] = 0;                                                // leave the state clean for the next pass / call
    if (threadIdx.x == 0) {
        st->prefix = prefix | (d << shift);
        st->mask = mask | (0xFFu << shift);
        st->k_rem = k_rem - above;
        st->ticket = 0;
    }
}

// elements equal to the threshold in a block (several levels can tie when the scale is 0 or the products are subnormal)
__device__ __forceinline__ uint32_t block_ties(uint64_t cnt, uint32_t valid, float s7, uint32_t t) {
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j <= 8; ++j) c += level_bits(s7, j) == t ? level_count(cnt, valid, j) : 0u;
    return c;
}

// each CTA owns one contiguous range of blocks (index order): its number of ties
__global__ void __launch_bounds__(kThrThreads)
k_thr4_count_ties(const uint64_t *__restrict__ levels, const float *__restrict__ scales, uint64_t n, uint64_t nblocks,
                  uint64_t blocks_per_cta, const ThrState *__restrict__ st, uint32_t *__restrict__ tie_count) {
    const uint32_t t = st->prefix;
    const uint64_t b0 = (uint64_t)blockIdx.x * blocks_per_cta, b1 = min(b0 + blocks_per_cta, nblocks);
    uint32_t c = 0;
    for (uint64_t b = b0 + threadIdx.x; b < b1; b += kThrThreads)
        c += block_ties(levels[b], (uint32_t)(n - b * 64 < 64 ? n - b * 64 : 64), __fdiv_rn(scales[b], 7.0f), t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    __shared__ uint32_t ws[kThrThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kThrThreads / 32; ++w) s += ws[w];
        tie_count[blockIdx.x] = s;
    }
}

// thread = block, the CTA walks its range in index order: levels above t stay, below t go, a nibble whose level equals t
// stays while fewer than keep_ties such elements precede it
template <int NT>
__device__ __forceinline__ void thr4_apply_range(uint4 *__restrict__ values, const uint64_t *__restrict__ levels, const float *__restrict__ scales,
                                                 uint64_t n, uint64_t b0, uint64_t b1, uint32_t t, uint64_t keep_ties, uint64_t first_rank) {
    __shared__ uint32_t ws[NT / 32];
    __shared__ uint64_t running;
    if (threadIdx.x == 0) running = first_rank;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t base = b0; base < b1; base += NT) {
        const uint64_t b = base + threadIdx.x;
        const bool live = b < b1;
        uint32_t valid = 0, ties = 0, keep_mask = 0, tie_mask = 0;           // bit j: level j survives / ties
        if (live) {
            valid = (uint32_t)(n - b * 64 < 64 ? n - b * 64 : 64);
            const uint64_t cnt = levels[b];
            const float s7 = __fdiv_rn(scales[b], 7.0f);
#pragma unroll
            for (int j = 0; j <= 8; ++j) {
                const uint32_t m = level_bits(s7, j);
                keep_mask |= (m > t ? 1u : 0u) << j;
                tie_mask |= (m == t ? 1u : 0u) << j;
                ties += m == t ? level_count(cnt, valid, j) : 0u;
            }
        }
        uint32_t incl = ties;                                                // exclusive scan of `ties` in thread (= index) order
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int x = 0; x < NT / 32; ++x) { const uint32_t v = ws[x]; if (x < warp) before += v; total += v; }
        uint64_t rank = running + before + (incl - ties);
        // Levels are monotone in j for an ordinary scale, so the survivors of a block are "every nibble with level >= jc" as
        // long as its ties are kept or dropped as a whole: eight nibbles per word are then filtered at once (byte-wise add
        // of 16 - jc to the levels spread over bytes: bit 4 of a byte = level >= jc). Blocks where a tie level is cut in
        // the middle, or whose levels do not order (zero / non-finite scale, subnormal products), walk their nibbles.
        uint32_t eff = keep_mask;
        bool fast = true;
        if (ties) {
            if (rank + ties <= keep_ties) eff |= tie_mask;
            else if (rank < keep_ties) fast = false;
        }
        const uint32_t jc = eff ? (uint32_t)__ffs(eff) - 1u : 9u;
        fast = fast && eff == ((0x1FFu >> jc) << jc);
        if (live && !(fast && jc == 0)) {                                   // jc == 0: everything stays, the block is not rewritten
            uint4 v[2] = {values[2 * b], values[2 * b + 1]};
            uint32_t *w = reinterpret_cast<uint32_t *>(v);
            if (fast) {
                const uint32_t add = (16u - jc) * 0x01010101u;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t a = nibble_abs8(w[i]);
                    const uint32_t ke = (((a & 0x0F0F0F0Fu) + add) >> 4) & 0x01010101u;          // low nibbles with level >= jc
                    const uint32_t ko = ((((a >> 4) & 0x0F0F0F0Fu) + add) >> 4) & 0x01010101u;   // high nibbles
                    w[i] &= ke * 0x0Fu + ko * 0xF0u;
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                    const uint32_t a = nibble_abs8(w[i]);
                    uint32_t out = w[i];
                    for (int e = 0; e < 8; ++e) {
                        if ((uint32_t)(8 * i + e) >= valid) continue;
                        const int sh = 4 * (e ^ 1);
                        const uint32_t lv = (a >> sh) & 0xFu;
                        bool keep = (keep_mask >> lv) & 1u;
                        if ((tie_mask >> lv) & 1u) { keep = rank < keep_ties; ++rank; }
                        if (!keep) out &= ~(0xFu << sh);
                    }
                    w[i] = out;
                }
            }
            values[2 * b] = v[0];
            values[2 * b + 1] = v[1];
        }
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThrThreads)
k_thr4_apply(uint4 *__restrict__ values, const uint64_t *__restrict__ levels, const float *__restrict__ scales, uint64_t n,
             uint64_t nblocks, uint64_t blocks_per_cta, const ThrState *__restrict__ st, const uint64_t *__restrict__ tie_base) {
    const uint64_t b0 = (uint64_t)blockIdx.x * blocks_per_cta, b1 = min(b0 + blocks_per_cta, nblocks);
    thr4_apply_range<kThrThreads>(values, levels, scales, n, b0, b1, st->prefix, st->k_rem, tie_base[blockIdx.x]);
}

// ---- FAST (n <= kThrSmallLimit, the IHT sizes): ONE CTA does the four digit passes and the ordered apply - one launch.
// The passes are bound by the instructions that rebuild a magnitude (~30 per element: nibble extract, int->float, IEEE
// divide / multiply), so the first pass leaves the magnitudes in an L2-resident scratch array and the other passes and the
// apply read them back (one 16-byte load per four elements).
template <int BITS>
__global__ void __launch_bounds__(kThrSmallThreads)
k_thr_small(uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint32_t nwords, uint64_t k,
            uint32_t *__restrict__ mag) {
    constexpr int E = ThrWord<BITS>::kElems;
    __shared__ uint32_t h[256];
    __shared__ uint64_t scratch[256];
    uint32_t prefix = 0, mask = 0;
    uint64_t k_rem = k;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) h[threadIdx.x] = 0;
        __syncthreads();
        if (shift == 24) {
            for (uint32_t i = threadIdx.x; i < nwords; i += kThrSmallThreads) {
                uint32_t m[E];
                ThrWord<BITS>::mags(values[i], scales[((uint64_t)i * E) >> 6], (uint64_t)i * E, n, m);
#pragma unroll
                for (int j = 0; j < E / 4; ++j)
                    reinterpret_cast<uint4 *>(mag)[(uint64_t)i * (E / 4) + j] = make_uint4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]);
#pragma unroll
                for (int e = 0; e < E; ++e) hist_add(h, m[e] >> 24, (uint64_t)i * E + e < n);
            }
        } else {
#pragma unroll 2
            for (uint32_t i = threadIdx.x; i < nwords; i += kThrSmallThreads) {
                uint32_t m[E];
                load_mags<E>(mag, i, m);
#pragma unroll
                for (int e = 0; e < E; ++e) hist_add(h, (m[e] >> shift) & 0xFFu, (uint64_t)i * E + e < n && (m[e] & mask) == prefix);
            }
        }
        __syncthreads();
        uint64_t above;
        const uint32_t d = pick_digit(h, k_rem, scratch, above);
        prefix |= d << shift;
        mask |= 0xFFu << shift;
        k_rem -= above;
        __syncthreads();
    }
    apply_range<BITS, kThrSmallThreads>(values, scales, n, 0, nwords, prefix, k_rem, 0, mag);
}

// ---- FAST (kThrSmallLimit < n <= kThrClusterLimit): the same single launch spread over a thread-block cluster. Every CTA
// histograms its contiguous slice of the words, the eight partial histograms are summed through distributed shared
// memory (each CTA reads its peers' bins directly), every CTA picks the digit redundantly - no global state, no host
// sync, two histogram buffers so that one cluster barrier per pass suffices. Tie ranks: the CTAs publish their tie
// counts in shared memory and each adds up the counts of the ranks before it.
template <int BITS>
__global__ void __launch_bounds__(kThrSmallThreads)
k_thr_cluster(uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint32_t nwords, uint64_t k,
              uint32_t *__restrict__ mag) {
    namespace cg = cooperative_groups;
    constexpr int E = ThrWord<BITS>::kElems;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank(), csize = cluster.num_blocks();
    __shared__ uint32_t h[2][256];
    __shared__ uint32_t hsum[256];
    __shared__ uint64_t scratch[256];
    __shared__ uint32_t my_ties;
    const uint32_t per = (nwords + csize - 1) / csize, w0 = min(rank * per, nwords), w1 = min(w0 + per, nwords);
    uint32_t prefix = 0, mask = 0;
    uint64_t k_rem = k;
    if (threadIdx.x < 256) { h[0][threadIdx.x] = 0; h[1][threadIdx.x] = 0; }
    if (threadIdx.x == 0) my_ties = 0;
    __syncthreads();
    int buf = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8, buf ^= 1) {
        uint32_t *hl = h[buf];
        if (shift == 24) {
            for (uint32_t i = w0 + threadIdx.x; i < w1; i += kThrSmallThreads) {
                uint32_t m[E];
                ThrWord<BITS>::mags(values[i], scales[((uint64_t)i * E) >> 6], (uint64_t)i * E, n, m);
#pragma unroll
                for (int j = 0; j < E / 4; ++j)
                    reinterpret_cast<uint4 *>(mag)[(uint64_t)i * (E / 4) + j] = make_uint4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]);
#pragma unroll
                for (int e = 0; e < E; ++e) hist_add(hl, m[e] >> 24, (uint64_t)i * E + e < n);
            }
        } else {
#pragma unroll 2
            for (uint32_t i = w0 + threadIdx.x; i < w1; i += kThrSmallThreads) {
                uint32_t m[E];
                load_mags<E>(mag, i, m);
#pragma unroll
                for (int e = 0; e < E; ++e) hist_add(hl, (m[e] >> shift) & 0xFFu, (uint64_t)i * E + e < n && (m[e] & mask) == prefix);
            }
        }
        cluster.sync();                                                  // every CTA's partial histogram is complete
        if (threadIdx.x < 256) {
            uint32_t sum = 0;
            for (uint32_t r = 0; r < csize; ++r) sum += cluster.map_shared_rank(&h[buf][0], r)[threadIdx.x];
            hsum[threadIdx.x] = sum;
            h[buf ^ 1][threadIdx.x] = 0;                                 // the peers finished reading it before this pass's barrier
        }
        __syncthreads();
        uint64_t above;
        const uint32_t d = pick_digit(hsum, k_rem, scratch, above);
        prefix |= d << shift;
        mask |= 0xFFu << shift;
        k_rem -= above;
        __syncthreads();
    }
    // ties at the threshold in this CTA's slice
    uint32_t c = 0;
    for (uint32_t i = w0 + threadIdx.x; i < w1; i += kThrSmallThreads) {
        uint32_t m[E];
        load_mags<E>(mag, i, m);
#pragma unroll
        for (int e = 0; e < E; ++e) c += ((uint64_t)i * E + e < n && m[e] == prefix) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&my_ties, c);
    cluster.sync();
    uint64_t first_rank = 0;
    for (uint32_t r = 0; r < rank; ++r) first_rank += *cluster.map_shared_rank(&my_ties, r);
    apply_range<BITS, kThrSmallThreads>(values, scales, n, w0, w1, prefix, k_rem, first_rank, mag);
    cluster.sync();                                                      // nobody leaves while a peer may still read its shared memory
}

// ---- EXACT: magnitudes + sign-extended bits in parallel, the heap walk by one thread, the mask applied in parallel --
template <int BITS>
__global__ void k_thr_prepare(const uint32_t *__restrict__ values, const float *__restrict__ scales, uint64_t n, uint64_t nwords,
                              float *__restrict__ mag, uint8_t *__restrict__ keep) {
    constexpr int E = ThrWord<BITS>::kElems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t m[E];
        ThrWord<BITS>::mags(values[i], scales[(i * E) >> 6], i * E, n, m);
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (i * E + e < n) { mag[i * E + e] = __uint_as_float(m[e]); keep[i * E + e] = 0; }
    }
}

struct HeapItem { float value; uint32_t idx; };
__device__ __forceinline__ bool heap_gt(const HeapItem &a, const HeapItem &b) { return a.value > b.value || isnan(a.value); }   // gt_idx_t, include/CloverBase.h:205-224

__device__ void heap_push(HeapItem *first, long hole, long top, HeapItem value) {          // libstdc++ __push_heap
    long parent = (hole - 1) / 2;
    while (hole > top && heap_gt(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
__device__ void heap_adjust(HeapItem *first, long hole, long len, HeapItem value) {        // libstdc++ __adjust_heap
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (heap_gt(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    heap_push(first, hole, top, value);
}

__global__ void k_thr_heap(const float *__restrict__ mag, uint64_t n, uint64_t k, HeapItem *__restrict__ heap, uint8_t *__restrict__ keep) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (uint64_t i = 0; i < k; ++i) heap[i] = HeapItem{mag[i], (uint32_t)i};
    if (k >= 2) {                                                                          // std::make_heap
        long parent = ((long)k - 2) / 2;
        for (;;) {
            const HeapItem v = heap[parent];
            heap_adjust(heap, parent, (long)k, v);
            if (parent == 0) break;
            parent--;
        }
    }
    for (uint64_t i = k; i < n; ++i) {
        const float value = mag[i];
        if (value > heap[0].value) {
            heap[0] = HeapItem{value, (uint32_t)i};
            uint32_t pos = 0;                                                              // min_heapify, include/CloverBase.h:226-249
            for (;;) {
                const uint32_t l = 2 * pos + 1, r = 2 * pos + 2;
                uint32_t smallest = pos;
                if (l < k && heap[l].value < heap[smallest].value) smallest = l;
                if (r < k && heap[r].value < heap[smallest].value) smallest = r;
                if (smallest == pos) break;
                const HeapItem tmp = heap[pos]; heap[pos] = heap[smallest]; heap[smallest] = tmp;
                pos = smallest;
            }
        }
    }
    for (uint64_t i = 0; i < k; ++i) keep[heap[i].idx] = 1;
}

template <int BITS>
__global__ void k_thr_apply_mask(uint32_t *__restrict__ values, uint64_t n, uint64_t nwords, const uint8_t *__restrict__ keep) {
    constexpr int E = ThrWord<BITS>::kElems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = values[i];
        uint32_t out = w;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (i * E + e < n && !keep[i * E + e]) out = ThrWord<BITS>::clear(out, e);
        if (out != w) values[i] = out;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct ThrWorkspace { void *p; size_t bytes; };
// workspace keyed by (device, stream): hist[] / ticket (the ThrState at its head) start clean and every pass leaves them clean
static int thr_workspace(cudaStream_t stream, size_t bytes, void **out) {
    return stream_scratch(kScratchThreshold, stream, bytes, sizeof(ThrState), out);
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <int BITS>
static int launch_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k, int mode, cudaStream_t stream) {
    constexpr int E = ThrWord<BITS>::kElems;
    if (n == 0 || k >= n) return CLOVER_OK;                          // k == n keeps everything; the reference requires k <= n
    const uint64_t nwords = (n + E - 1) / E;
    uint32_t *v32 = reinterpret_cast<uint32_t *>(values);
    const uint64_t cap = (uint64_t)sm_count() * 8;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nwords + kThrThreads - 1) / kThrThreads, cap));
    if (k == 0) {                                                    // nothing survives (pad nibbles are zero already)
        CLOVER_CUDA_CHECK(cudaMemsetAsync(values, 0, BITS == 4 ? (n + 1) / 2 : n, stream));
        return CLOVER_OK;
    }
    const bool exact = mode == CLOVER_THRESHOLD_EXACT || (mode == CLOVER_THRESHOLD_AUTO && n <= kThrExactLimit);
    if (exact) {
        CLOVER_REQUIRE(n <= 0xFFFFFFFFull, CLOVER_ERR_UNSUPPORTED, "EXACT threshold is limited to 2^32 - 1 elements");
        const size_t off_mag = align_up(sizeof(ThrState), 256);          // the FAST path's state keeps the head of the workspace
        const size_t off_heap = off_mag + align_up(n * sizeof(float), 256), off_keep = off_heap + align_up(k * sizeof(HeapItem), 256);
        void *ws = nullptr;
        int rc = thr_workspace(stream, off_keep + align_up(n, 256), &ws);
        if (rc != CLOVER_OK) return rc;
        float *mag = reinterpret_cast<float *>(static_cast<uint8_t *>(ws) + off_mag);
        HeapItem *heap = reinterpret_cast<HeapItem *>(static_cast<uint8_t *>(ws) + off_heap);
        uint8_t *keep = static_cast<uint8_t *>(ws) + off_keep;
        k_thr_prepare<BITS><<<grid, kThrThreads, 0, stream>>>(v32, scales, n, nwords, mag, keep);
        k_thr_heap<<<1, 32, 0, stream>>>(mag, n, k, heap, keep);
        k_thr_apply_mask<BITS><<<grid, kThrThreads, 0, stream>>>(v32, n, nwords, keep);
        count_launch(3);
        return launch_status("k_thr_heap");
    }
    if (n <= kThrSmallLimit) {
        const size_t off_mag = align_up(sizeof(ThrState), 256);
        void *ws = nullptr;
        int rc = thr_workspace(stream, off_mag + nwords * E * sizeof(uint32_t), &ws);
        if (rc != CLOVER_OK) return rc;
        k_thr_small<BITS><<<1, kThrSmallThreads, 0, stream>>>(v32, scales, n, (uint32_t)nwords, k,
                                                              reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + off_mag));
        count_launch();
        return launch_status("k_thr_small");
    }
    if (n <= kThrClusterLimit) {
        const size_t off_mag = align_up(sizeof(ThrState), 256);
        void *ws = nullptr;
        int rc = thr_workspace(stream, off_mag + nwords * E * sizeof(uint32_t), &ws);
        if (rc != CLOVER_OK) return rc;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kThrClusterSize);
        cfg.blockDim = dim3(kThrSmallThreads);
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kThrClusterSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CLOVER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_thr_cluster<BITS>, v32, scales, n, (uint32_t)nwords, k,
                                             reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + off_mag)));
        count_launch();
        return launch_status("k_thr_cluster");
    }
    if (BITS == 4 && (reinterpret_cast<uintptr_t>(values) & 15u) == 0) {
        // selection over per-block level counts (see k_thr4_levels); n_pad is a multiple of 128, so whole blocks are readable
        const uint64_t nblocks = (n + 63) / 64;
        const unsigned bgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nblocks + kThrThreads - 1) / kThrThreads, cap));
        const uint64_t blocks_per_cta = (nblocks + bgrid - 1) / bgrid;
        const size_t off_cnt = align_up(sizeof(ThrState), 256), off_base = off_cnt + align_up(sizeof(uint32_t) * bgrid, 256);
        const size_t off_lev = off_base + align_up(sizeof(uint64_t) * bgrid, 256);
        void *ws = nullptr;
        int rc = thr_workspace(stream, off_lev + nblocks * sizeof(uint64_t), &ws);
        if (rc != CLOVER_OK) return rc;
        ThrState *st = static_cast<ThrState *>(ws);
        uint32_t *tie_count = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + off_cnt);
        uint64_t *tie_base = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(ws) + off_base);
        uint64_t *levels = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(ws) + off_lev);
        uint4 *v128 = reinterpret_cast<uint4 *>(values);
        k_thr4_levels<<<bgrid, kThrThreads, 0, stream>>>(v128, n, nblocks, levels);
        k_thr4_hist<true><<<bgrid, kThrThreads, 0, stream>>>(levels, scales, n, nblocks, 24, k, st);
        for (int shift = 16; shift >= 0; shift -= 8)
            k_thr4_hist<false><<<bgrid, kThrThreads, 0, stream>>>(levels, scales, n, nblocks, shift, k, st);
        k_thr4_count_ties<<<bgrid, kThrThreads, 0, stream>>>(levels, scales, n, nblocks, blocks_per_cta, st, tie_count);
        k_thr_scan<<<1, 256, 0, stream>>>(tie_count, tie_base, (int)bgrid);
        k_thr4_apply<<<bgrid, kThrThreads, 0, stream>>>(v128, levels, scales, n, nblocks, blocks_per_cta, st, tie_base);
        count_launch(8);
        return launch_status("k_thr4_apply");
    }
    const uint64_t words_per_cta = (nwords + grid - 1) / grid;
    const size_t off_cnt = align_up(sizeof(ThrState), 256), off_base = off_cnt + align_up(sizeof(uint32_t) * grid, 256);
    void *ws = nullptr;
    int rc = thr_workspace(stream, off_base + sizeof(uint64_t) * grid, &ws);
    if (rc != CLOVER_OK) return rc;
    ThrState *st = static_cast<ThrState *>(ws);
    uint32_t *tie_count = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + off_cnt);
    uint64_t *tie_base = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(ws) + off_base);
    k_thr_hist<BITS, true><<<grid, kThrThreads, 0, stream>>>(v32, scales, n, nwords, 24, k, st);
    for (int shift = 16; shift >= 0; shift -= 8)
        k_thr_hist<BITS, false><<<grid, kThrThreads, 0, stream>>>(v32, scales, n, nwords, shift, k, st);
    k_thr_count_ties<BITS><<<grid, kThrThreads, 0, stream>>>(v32, scales, n, nwords, words_per_cta, st, tie_count);
    k_thr_scan<<<1, 256, 0, stream>>>(tie_count, tie_base, (int)grid);
    k_thr_apply<BITS><<<grid, kThrThreads, 0, stream>>>(v32, scales, n, nwords, words_per_cta, st, tie_base);
    count_launch(7);
    return launch_status("k_thr_apply");
}

}  // namespace clover

using namespace clover;

extern "C" {

uint64_t clover_threshold_exact_limit(void) { return kThrExactLimit; }

int clover_v4_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k, int mode, void *stream) {
    CLOVER_REQUIRE(values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(mode >= 0 && mode <= 2, CLOVER_ERR_INVALID, "bad threshold mode");
    return launch_threshold<4>(values, scales, n, k, mode, (cudaStream_t)stream);
}
int clover_v8_threshold(int8_t *values, const float *scales, uint64_t n, uint64_t k, int mode, void *stream) {
    CLOVER_REQUIRE(values && scales, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(mode >= 0 && mode <= 2, CLOVER_ERR_INVALID, "bad threshold mode");
    return launch_threshold<8>(values, scales, n, k, mode, (cudaStream_t)stream);
}

}  // extern "C"
