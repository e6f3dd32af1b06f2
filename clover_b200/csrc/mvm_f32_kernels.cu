// mvm_f32_kernels.cu - mvm with fp32 vectors: CloverMatrix4::mvm(V32,V32) and CloverMatrix8::mvm(V32,V32).
//
//   4-bit  include/CloverMatrix4.h:1451-1547:  s = su[b] / 7.0f;   f = float(q) * s;   acc = fma(x, f, acc)
//   8-bit  include/CloverMatrix8.h:558-661:    s = su[b] / 127.0f; t = x * s;          acc = fma(t, float(q), acc)
//
// Both run 32 fp32 chains per row: accumulator k = 0..3, AVX lane l = 0..7; per block chain (k, l) takes element
// 8k+l and then element 32+8k+l; the row result is ((acc0+acc1)+(acc2+acc3)) folded by the hadd tree of
// include/CloverBase.h:149-157. Every element costs an int->float, a rounded multiply and an fma in that order, so
// these kernels are bound by the CUDA cores' issue rate, not by HBM; the design minimises instructions per element:
//
//   * thread = half of one matrix row's chains (accumulators 2kh, 2kh+1 = 16 of the 32 chains, 8 packed f32x2
//     accumulators), warp = 16 consecutive rows (one work item): no shuffles anywhere in the loop, one exchange
//     between the two threads of a row at its end; 16-row items give 2048 items at 32768 rows = 14 warps per SM
//     (32-row items: 7 warps per SM issued in 61 % of the cycles, ncu r02b);
//   * the matrix streams through per-warp TMA rings: cp.async.bulk.tensor boxes of 16 rows x 128 bytes with the
//     128-byte swizzle (a quarter warp's 16-byte loads fall into 8 different 16-byte bank groups), the matching slice
//     of x by a 1D bulk copy into the same stage; the warp's own lane 0 refills a stage as soon as the warp has
//     consumed it (no producer warp, no empty-barriers);
//   * int -> float without I2F: the biased nibble / byte is spliced under the bits of 1.5 * 2^23 by one PRMT and the
//     bias removed by a packed add; multiply and fma are packed too (add/mul/fma.rn.f32x2 round each half like the
//     scalar instruction) - 2.9 (4-bit) / 3.1 (8-bit) instructions per element;
//   * x is read from shared memory as warp-wide broadcasts (every row needs the same 64 floats per block).
//
// k_mvm_f32_simple (one warp per row, plain loads) is the same arithmetic for unaligned operands and the cross-check.
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "async_copy.cuh"
#include "common.cuh"
#include "runtime.cuh"

namespace clover {

constexpr int kF32Rows = 16;            // rows per work item (= per warp)
constexpr int kF32Warps = 16;           // warps (= concurrent work items) per CTA
constexpr int kF32Stages = 4;           // ring depth per warp
constexpr int kF32StageBytes = 3072;    // 16 rows x 128 B + up to 256 floats of x, a multiple of 1024 (swizzle atom)

constexpr int kF32SmemBytes = kF32Warps * kF32Stages * kF32StageBytes + kF32Warps * kF32Stages * 8 + 1024 /* alignment slack */;

__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) { return pack2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void fma2(uint64_t &acc, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ uint32_t prmt_f32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ float lo_f(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }

constexpr uint32_t kMagic = 0x4B400000u;                 // 1.5 * 2^23: as_float(kMagic | u) = 12582912 + u for u < 2^22

// Constants that must live in REGISTERS: SASS PRMT / LOP3 take one immediate only. With the magic bits as an immediate
// ptxas keeps the PRMT selectors in uniform registers and re-materialises each one with an extra IMAD.U32 per PRMT;
// with two immediates a mask-and-bias costs two LOP3. `zero` is a kernel argument the compiler cannot fold (always 0 at run time).
struct RegConsts { uint32_t magic, m0f, c08, c80; };
__device__ __forceinline__ RegConsts reg_consts(uint32_t zero) {        // zero: a kernel ARGUMENT that is always 0
    return {kMagic | zero, 0x0F0F0F0Fu | zero, 0x08080808u | zero, 0x80808080u | zero};
}
__device__ __forceinline__ uint32_t and_xor(uint32_t a, uint32_t m, uint32_t c) {      // (a & m) ^ c in one LOP3
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
    return d;
}
// The per-pair arithmetic as ONE asm block each, so that ptxas allocates the two PRMT results as an aligned register pair
// (built from separate C++ values the pair cost an extra IMAD.MOV per element, 11 % of all instructions - ncu r02e).
//   pair4: acc += x2 * ((splice(hi, P) | splice(lo, P)) - bias) * s      f = float(q) * s (:1520-1527), acc = fma(x, f, acc) (:1529-1537)
//   pair8: acc += (x2 * s) * ((splice(u, P) | splice(u, P + 1)) - bias)  t = x * s (CloverMatrix8.h:632-639), acc = fma(t, float(q), acc) (:641-649)
template <int P>
__device__ __forceinline__ void pair4(uint64_t &acc, uint64_t x2, uint32_t hi, uint32_t lo, uint32_t magic, uint64_t neg, uint64_t ss) {
    asm("{\n\t.reg .b32 a, b;\n\t.reg .b64 m;\n\t"
        "prmt.b32 a, %2, %4, %7;\n\t"
        "prmt.b32 b, %3, %4, %7;\n\t"
        "mov.b64 m, {a, b};\n\t"
        "add.rn.f32x2 m, m, %5;\n\t"
        "mul.rn.f32x2 m, m, %6;\n\t"
        "fma.rn.f32x2 %0, %1, m, %0;\n\t}"
        : "+l"(acc) : "l"(x2), "r"(hi), "r"(lo), "r"(magic), "l"(neg), "l"(ss), "n"(0x7650 + P));
}
// (t2 = x2 * s is the same for all 16 rows of a work item - one 64x64 tile scale per block - so the warp multiplies the
//  stage's x slice ONCE, in place, before the block steps: one mul.f32x2 per 32 elements instead of one per 2)
template <int P>
__device__ __forceinline__ void pair8(uint64_t &acc, uint64_t t2, uint32_t u, uint32_t magic, uint64_t neg) {
    asm("{\n\t.reg .b32 a, b;\n\t.reg .b64 m;\n\t"
        "prmt.b32 a, %2, %3, %5;\n\t"
        "prmt.b32 b, %2, %3, %6;\n\t"
        "mov.b64 m, {a, b};\n\t"
        "add.rn.f32x2 m, m, %4;\n\t"
        "fma.rn.f32x2 %0, %1, m, %0;\n\t}"
        : "+l"(acc) : "l"(t2), "r"(u), "r"(magic), "l"(neg), "n"(0x7650 + P), "n"(0x7651 + P));
}

// x as packed f32x2 operands straight from shared memory (immediate offsets: no address arithmetic per load)
template <int OFF> __device__ __forceinline__ void lds_x4(uint32_t base, uint64_t &a, uint64_t &b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2+%3];" : "=l"(a), "=l"(b) : "r"(base), "n"(OFF));
}

// one 32-bit word of the 4-bit matrix (elements e..e+7 of a row: byte i = element 2i in the HIGH nibble, 2i+1 in the
// low one) against x[e..e+7] at shared address xaddr + XOFF
template <int XOFF>
__device__ __forceinline__ void word4(uint32_t w, uint32_t xaddr, uint64_t ss, uint64_t neg, const RegConsts &rc, uint64_t *acc) {
    const uint32_t lo = and_xor(w, rc.m0f, rc.c08);                  // byte p: q + 8 of element 2p+1
    const uint32_t hi = and_xor(w >> 4, rc.m0f, rc.c08);             // byte p: q + 8 of element 2p
    uint64_t x0, x1, x2, x3;
    lds_x4<XOFF>(xaddr, x0, x1);
    lds_x4<XOFF + 16>(xaddr, x2, x3);
    pair4<0>(acc[0], x0, hi, lo, rc.magic, neg, ss);
    pair4<1>(acc[1], x1, hi, lo, rc.magic, neg, ss);
    pair4<2>(acc[2], x2, hi, lo, rc.magic, neg, ss);
    pair4<3>(acc[3], x3, hi, lo, rc.magic, neg, ss);
}
// one 32-bit word of the 8-bit matrix (elements e..e+3) against t[e..e+3] = x * s at shared address xaddr + XOFF
template <int XOFF>
__device__ __forceinline__ void word8(uint32_t w, uint32_t xaddr, uint64_t neg, const RegConsts &rc, uint64_t *acc) {
    const uint32_t u = w ^ rc.c80;                                    // byte i: q + 128
    uint64_t x0, x1;
    lds_x4<XOFF>(xaddr, x0, x1);
    pair8<0>(acc[0], x0, u, rc.magic, neg);
    pair8<2>(acc[1], x1, u, rc.magic, neg);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

template <int MBITS>
__global__ void __launch_bounds__(kF32Warps * 32, 1)
k_mvm_f32_ring(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ scales, uint64_t rows, uint64_t cols,
               const float *__restrict__ x, float *__restrict__ y, uint32_t zero) {
    extern __shared__ uint8_t smem_raw_f32[];
    constexpr int kBPC = MBITS == 4 ? 4 : 2;                          // blocks of 64 columns per 128-byte chunk
    constexpr float kQ = MBITS == 4 ? 7.0f : 127.0f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // thread = (row r of the 16-row work item, accumulator pair kh: accumulators 2kh, 2kh+1 = elements 16kh .. 16kh+15 of
    // each half block). The mapping keeps every shared-memory load conflict-free under the 128-byte swizzle: the 4-bit
    // kernel loads 8 bytes per half block (a half warp = 8 rows x 2 kh), the 8-bit kernel 16 bytes (a quarter warp = 8 rows).
    const int r = MBITS == 4 ? lane >> 1 : lane & 15, kh = MBITS == 4 ? lane & 1 : lane >> 4;
    const uint32_t hb = (uint32_t)(cols >> 6), nitems = (uint32_t)(rows / kF32Rows);
    const uint32_t nchunks = (hb + kBPC - 1) / kBPC;
    const uint32_t wstride = gridDim.x * kF32Warps;
    const uint32_t first = (uint32_t)warp * gridDim.x + blockIdx.x;      // warp-major: every SM gets the same number of busy warps
    if (first >= nitems) return;

    // this warp's ring: stages at base + s * kF32StageBytes (1024-aligned), barriers behind all rings
    const uint32_t smem0 = (smem_u32(smem_raw_f32) + 1023u) & ~1023u;
    const uint32_t ring = smem0 + (uint32_t)warp * (kF32Stages * kF32StageBytes);
    const uint32_t bars = smem0 + kF32Warps * kF32Stages * kF32StageBytes + (uint32_t)warp * (8 * kF32Stages);
    if (lane == 0) {
        for (int s = 0; s < kF32Stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * s), "r"(1) : "memory");
        mbar_fence_init();
        tma_prefetch_descriptor(&tmap);
    }
    __syncwarp();

    // producer state (lane 0): the next (item, chunk) to request and its stage
    uint32_t p_item = first, p_c = 0, p_s = 0;
    auto issue = [&]() {                                               // lane 0 only; no-op when the warp's work is exhausted
        if (p_item >= nitems) return;
        const uint32_t nb = min((uint32_t)kBPC, hb - p_c * kBPC);
        const uint32_t dst = ring + p_s * kF32StageBytes, bar = bars + 8 * p_s;
        // (the warp's own reads of this stage are ordered before the refill by the __syncwarp() in front of this call - the
        //  generic-proxy hand-over every TMA consumer/producer ring uses; no cross-proxy fence is needed for a write AFTER reads)
        mbar_arrive_expect_tx_a(bar, kF32Rows * 128 + nb * 256);
        tma_load_2d_a(dst, &tmap, (int)(p_c * 128), (int)(p_item * kF32Rows), bar);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + kF32Rows * 128u), "l"(x + (uint64_t)p_c * kBPC * 64), "r"(nb * 256u), "r"(bar) : "memory");
        if (++p_s == kF32Stages) p_s = 0;
        if (++p_c == nchunks) { p_c = 0; p_item += wstride; }
    };
    if (lane == 0)
        for (int i = 0; i < kF32Stages; ++i) issue();

    // Block scales: lane L holds s = su[b] / 7.0f (:1488) resp. / 127.0f (CloverMatrix8.h:587) of block 32g + L of the current
    // row - ONE load and ONE IEEE divide per lane and 32 blocks (kCPG chunks); the raw value of the next group is requested
    // a group ahead. Per chunk the warp only shuffles.
    constexpr uint32_t kCPG = 32 / kBPC;                               // chunks per group of 32 blocks
    auto load_scale = [&](uint32_t item, uint32_t g) -> float {
        if (item >= nitems) return 1.0f;
        const uint32_t b = min(g * 32u + (uint32_t)lane, hb - 1);
        return __ldg(scales + (uint64_t)(item / (64 / kF32Rows)) * hb + b);
    };

    const uint64_t neg = MBITS == 4 ? pack2f(-12582920.0f, -12582920.0f) : pack2f(-12583040.0f, -12583040.0f);   // -(magic + bias)
    const RegConsts rc = reg_consts(zero);
    const uint32_t rsw = (uint32_t)(r & 7), rowoff = (uint32_t)r * 128u;
    uint32_t s = 0, phase = 0;
    float sraw = load_scale(first, 0);
    for (uint32_t item = first; item < nitems; item += wstride) {
        uint64_t acc[2][4];                                            // accumulators 2kh, 2kh+1; pair p = AVX lanes 2p, 2p+1
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[k][p] = 0ull;
        float sdiv = 0.f;
        for (uint32_t c = 0; c < nchunks; ++c) {
            if (c % kCPG == 0) {
                sdiv = __fdiv_rn(sraw, kQ);
                sraw = (c + kCPG) * kBPC < hb ? load_scale(item, c / kCPG + 1) : load_scale(item + wstride, 0);
            }
            mbar_wait_a(bars + 8 * s, phase);
            const uint32_t st = ring + s * kF32StageBytes, rowp = st + rowoff, xs = st + kF32Rows * 128u + 64u * (uint32_t)kh;
            const int nb = (int)min((uint32_t)kBPC, hb - c * kBPC);
            // the block scales of the chunk, broadcast once (lane j of every group of kBPC lanes holds block j's)
            float sjs[4];
#pragma unroll
            for (int j = 0; j < kBPC; ++j) sjs[j] = __shfl_sync(0xFFFFFFFFu, sdiv, (int)(c % kCPG) * kBPC + j);
            if (MBITS == 8) {
                // t = x * s (CloverMatrix8.h:632-639) for the chunk's 128 elements, in place: lane L owns floats 4L .. 4L+3
                // (block L / 16). Padding blocks beyond hb hold stale bytes - they are multiplied but never read.
                const uint32_t ta = st + kF32Rows * 128u + 16u * (uint32_t)lane;
                const float sj = lane < 16 ? sjs[0] : sjs[1];
                const uint64_t s2 = pack2f(sj, sj);
                uint64_t a, b;
                asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(ta));
                a = mul2(a, s2); b = mul2(b, s2);
                asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(ta), "l"(a), "l"(b) : "memory");
                __syncwarp();
            }
            auto block_step = [&](auto J) {
                constexpr int j = decltype(J)::value;
                if (j < kBPC) {
                    const uint64_t ss = pack2f(sjs[j], sjs[j]);
                    // x[64j + 16kh ...] of this stage; the second half block is 128 bytes further
                    if (MBITS == 4) {
                        const uint2 w0 = lds64(rowp + ((((uint32_t)(2 * j)) ^ rsw) << 4) + 8u * (uint32_t)kh);        // elements 16kh .. 16kh+15
                        const uint2 w1 = lds64(rowp + ((((uint32_t)(2 * j + 1)) ^ rsw) << 4) + 8u * (uint32_t)kh);    // elements 32+16kh ..
                        word4<256 * j>(w0.x, xs, ss, neg, rc, acc[0]);               // element 8k+l, k = 2kh
                        word4<256 * j + 32>(w0.y, xs, ss, neg, rc, acc[1]);          //               k = 2kh+1
                        word4<256 * j + 128>(w1.x, xs, ss, neg, rc, acc[0]);         // element 32+8k+l
                        word4<256 * j + 160>(w1.y, xs, ss, neg, rc, acc[1]);
                    } else {
                        {   // half 0: this thread's 16 bytes = 16-byte chunk 4j + kh
                            const uint4 w = lds128(rowp + ((((uint32_t)(4 * j) + (uint32_t)kh) ^ rsw) << 4));
                            word8<256 * j>(w.x, xs, neg, rc, &acc[0][0]);          // accumulator 2kh, lanes 0..3
                            word8<256 * j + 16>(w.y, xs, neg, rc, &acc[0][2]);     //                   lanes 4..7
                            word8<256 * j + 32>(w.z, xs, neg, rc, &acc[1][0]);     // accumulator 2kh+1
                            word8<256 * j + 48>(w.w, xs, neg, rc, &acc[1][2]);
                        }
                        {   // half 1: chunk 4j + 2 + kh, x 128 bytes further
                            const uint4 w = lds128(rowp + ((((uint32_t)(4 * j + 2) + (uint32_t)kh) ^ rsw) << 4));
                            word8<256 * j + 128>(w.x, xs, neg, rc, &acc[0][0]);
                            word8<256 * j + 144>(w.y, xs, neg, rc, &acc[0][2]);
                            word8<256 * j + 160>(w.z, xs, neg, rc, &acc[1][0]);
                            word8<256 * j + 176>(w.w, xs, neg, rc, &acc[1][2]);
                        }
                    }
                }
            };
            // nb is warp-uniform. A full chunk (every chunk but possibly the last of a row) runs without any guard, so that
            // the accumulators stay in place; the partial chunk at a row end takes the guarded path.
            if (nb == kBPC) {
                block_step(std::integral_constant<int, 0>{});
                block_step(std::integral_constant<int, 1>{});
                block_step(std::integral_constant<int, 2>{});
                block_step(std::integral_constant<int, 3>{});
            } else {
                if (nb > 0) block_step(std::integral_constant<int, 0>{});
                if (nb > 1) block_step(std::integral_constant<int, 1>{});
                if (nb > 2) block_step(std::integral_constant<int, 2>{});
            }
            __syncwarp();
            if (lane == 0) {
                if (MBITS == 8) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp WROTE the stage (t = x * s) before the async-proxy refill
                issue();
            }
            if (++s == kF32Stages) { s = 0; phase ^= 1; }
        }
        // (acc_1 + acc_2) resp. (acc_3 + acc_4) in this thread, their sum across the thread pair, then the hadd tree
        // ((a4+a0)+(a6+a2))+((a5+a1)+(a7+a3)) (CloverBase.h:149-157)
        float t[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            t[2 * p] = __fadd_rn(lo_f(acc[0][p]), lo_f(acc[1][p]));
            t[2 * p + 1] = __fadd_rn(hi_f(acc[0][p]), hi_f(acc[1][p]));
        }
#pragma unroll
        for (int l = 0; l < 8; ++l) t[l] = __fadd_rn(t[l], __shfl_xor_sync(0xFFFFFFFFu, t[l], MBITS == 4 ? 1 : 16));   // fp32 add commutes bit for bit
        if (kh == 0)
            y[(uint64_t)item * kF32Rows + r] = __fadd_rn(__fadd_rn(__fadd_rn(t[4], t[0]), __fadd_rn(t[6], t[2])),
                                                         __fadd_rn(__fadd_rn(t[5], t[1]), __fadd_rn(t[7], t[3])));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same arithmetic with 8-row work items: thread = (row, ONE accumulator k: elements 8k..8k+7 of each half block, four
// packed f32x2 accumulators), 28 warps per CTA, and TWO 128-byte chunks per ring stage. What the 16-row kernel lacks is
// warps: 2048 items at 32768 rows are 14 per SM, 3.5 per scheduler, 1.5 of them eligible (ncu r02t), so the loop issues in
// 69 % of the cycles. Halving the item doubles the warps; a first cut with one chunk per stage lost more than that gained
// (228 vs 198 us: the per-stage work of a warp - barrier wait, scale shuffles, refill by lane 0 - was paid twice as often),
// so a stage now holds two chunks and the per-element overhead is that of the 16-row kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kR8Rows = 8, kR8Warps = 28, kR8Stages = 2, kR8Chunks = 2;
constexpr int kR8StageBytes = kR8Chunks * kR8Rows * 128 + kR8Chunks * 1024;     // 2 x (8 rows x 128 B) + up to 512 floats of x = 4096
constexpr int kR8SmemBytes = kR8Warps * kR8Stages * kR8StageBytes + kR8Warps * kR8Stages * 8 + 1024 /* alignment slack */;

template <int MBITS>
__global__ void __launch_bounds__(kR8Warps * 32, 1)
k_mvm_f32_ring8(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ scales, uint64_t rows, uint64_t cols,
                const float *__restrict__ x, float *__restrict__ y, uint32_t zero) {
    extern __shared__ uint8_t smem_raw_f32[];
    constexpr int kBPC = MBITS == 4 ? 4 : 2;                          // blocks of 64 columns per 128-byte chunk
    constexpr int kBPS = kR8Chunks * kBPC;                            // blocks per stage
    constexpr float kQ = MBITS == 4 ? 7.0f : 127.0f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // lane -> (row r, accumulator k). 4-bit: a warp-wide 4-byte load covers 8 rows x one 16-byte chunk; 8-bit: a half warp's
    // 8-byte loads cover 8 rows x one 16-byte chunk (k >> 1 fixed per half warp) - conflict-free under the 128-byte swizzle.
    const int r = MBITS == 4 ? lane >> 2 : (lane >> 1) & 7, k = MBITS == 4 ? lane & 3 : 2 * (lane >> 4) + (lane & 1);
    const uint32_t hb = (uint32_t)(cols >> 6), nitems = (uint32_t)(rows / kR8Rows);
    const uint32_t nchunks = (hb + kBPC - 1) / kBPC, nst = (nchunks + kR8Chunks - 1) / kR8Chunks;      // stages per item
    const uint32_t wstride = gridDim.x * kR8Warps;
    const uint32_t first = (uint32_t)warp * gridDim.x + blockIdx.x;      // warp-major: every SM gets the same number of busy warps
    if (first >= nitems) return;

    const uint32_t smem0 = (smem_u32(smem_raw_f32) + 1023u) & ~1023u;
    const uint32_t ring = smem0 + (uint32_t)warp * (kR8Stages * kR8StageBytes);
    const uint32_t bars = smem0 + kR8Warps * kR8Stages * kR8StageBytes + (uint32_t)warp * (8 * kR8Stages);
    if (lane == 0) {
        for (int s = 0; s < kR8Stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * s), "r"(1) : "memory");
        mbar_fence_init();
        tma_prefetch_descriptor(&tmap);
    }
    __syncwarp();

    // producer state (lane 0): the next (item, stage-of-item) to request and its ring slot
    uint32_t p_item = first, p_t = 0, p_s = 0;
    auto issue = [&]() {                                               // lane 0 only; no-op when the warp's work is exhausted
        if (p_item >= nitems) return;
        const uint32_t c0 = p_t * kR8Chunks;
        const uint32_t nblk = min((uint32_t)kBPS, hb - c0 * kBPC);                  // blocks of this stage (x floats = 64 nblk)
        const bool two = c0 + 1 < nchunks;
        const uint32_t dst = ring + p_s * kR8StageBytes, bar = bars + 8 * p_s;
        mbar_arrive_expect_tx_a(bar, kR8Rows * 128 * (two ? 2u : 1u) + nblk * 256);
        tma_load_2d_a(dst, &tmap, (int)(c0 * 128), (int)(p_item * kR8Rows), bar);
        if (two) tma_load_2d_a(dst + kR8Rows * 128, &tmap, (int)((c0 + 1) * 128), (int)(p_item * kR8Rows), bar);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + kR8Chunks * kR8Rows * 128u), "l"(x + (uint64_t)c0 * kBPC * 64), "r"(nblk * 256u), "r"(bar) : "memory");
        p_s ^= 1u;
        if (++p_t == nst) { p_t = 0; p_item += wstride; }
    };
    if (lane == 0) { issue(); issue(); }

    // block scales: lane L holds s = su[b] / Q of block 32g + L of the current row block; a group of 32 blocks = kSPG stages
    constexpr uint32_t kSPG = 32 / kBPS;
    auto load_scale = [&](uint32_t item, uint32_t g) -> float {
        if (item >= nitems) return 1.0f;
        const uint32_t b = min(g * 32u + (uint32_t)lane, hb - 1);
        return __ldg(scales + (uint64_t)(item / (64 / kR8Rows)) * hb + b);
    };

    const uint64_t neg = MBITS == 4 ? pack2f(-12582920.0f, -12582920.0f) : pack2f(-12583040.0f, -12583040.0f);   // -(magic + bias)
    const RegConsts rc = reg_consts(zero);
    const uint32_t rsw = (uint32_t)(r & 7), rowoff = (uint32_t)r * 128u;
    uint32_t s = 0, phase = 0;
    float sraw = load_scale(first, 0);
    for (uint32_t item = first; item < nitems; item += wstride) {
        uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};                    // accumulator k; pair p = AVX lanes 2p, 2p+1
        float sdiv = 0.f;
        for (uint32_t t = 0; t < nst; ++t) {
            if (t % kSPG == 0) {
                sdiv = __fdiv_rn(sraw, kQ);
                sraw = (t + kSPG) * kBPS < hb ? load_scale(item, t / kSPG + 1) : load_scale(item + wstride, 0);
            }
            mbar_wait_a(bars + 8 * s, phase);
            const uint32_t st = ring + s * kR8StageBytes, xbase = st + kR8Chunks * kR8Rows * 128u;
            const int nblk = (int)min((uint32_t)kBPS, hb - t * kBPS);
            float sjs[kBPS];
#pragma unroll
            for (int j = 0; j < kBPS; ++j) sjs[j] = __shfl_sync(0xFFFFFFFFu, sdiv, (int)(t % kSPG) * kBPS + j);
            if (MBITS == 8) {
                // t = x * s (CloverMatrix8.h:632-639) for the stage's 4 blocks = 256 floats, in place: lane L owns floats 8L .. 8L+7
                const uint32_t ta = xbase + 32u * (uint32_t)lane;
                const float sj = __shfl_sync(0xFFFFFFFFu, sdiv, (int)(t % kSPG) * kBPS + ((lane >> 3) & (kBPS - 1)));
                const uint64_t s2 = pack2f(sj, sj);
                uint64_t a, b, c, d;
                asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(ta));
                asm volatile("ld.shared.v2.b64 {%0,%1}, [%2+16];" : "=l"(c), "=l"(d) : "r"(ta));
                a = mul2(a, s2); b = mul2(b, s2); c = mul2(c, s2); d = mul2(d, s2);
                asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(ta), "l"(a), "l"(b) : "memory");
                asm volatile("st.shared.v2.b64 [%0+16], {%1,%2};" ::"r"(ta), "l"(c), "l"(d) : "memory");
                __syncwarp();
            }
            // block jj of the stage = chunk jj / kBPC, block jj % kBPC of the chunk; its x (or t) starts 256 jj bytes into the x area
            auto block_step = [&](auto JJ) {
                constexpr int jj = decltype(JJ)::value;
                if (jj < kBPS) {
                    constexpr int h = jj / kBPC, j = jj % kBPC;
                    const uint64_t ss = pack2f(sjs[jj], sjs[jj]);
                    const uint32_t rowp = st + (uint32_t)h * (kR8Rows * 128u) + rowoff, xs = xbase + 32u * (uint32_t)k;
                    if (MBITS == 4) {
                        const uint32_t w0 = lds32(rowp + ((((uint32_t)(2 * j)) ^ rsw) << 4) + 4u * (uint32_t)k);        // elements 8k .. 8k+7
                        const uint32_t w1 = lds32(rowp + ((((uint32_t)(2 * j + 1)) ^ rsw) << 4) + 4u * (uint32_t)k);    // elements 32+8k ..
                        word4<256 * jj>(w0, xs, ss, neg, rc, acc);
                        word4<256 * jj + 128>(w1, xs, ss, neg, rc, acc);
                    } else {
                        const uint2 w0 = lds64(rowp + ((((uint32_t)(4 * j) + (uint32_t)(k >> 1)) ^ rsw) << 4) + 8u * (uint32_t)(k & 1));
                        const uint2 w1 = lds64(rowp + ((((uint32_t)(4 * j + 2) + (uint32_t)(k >> 1)) ^ rsw) << 4) + 8u * (uint32_t)(k & 1));
                        word8<256 * jj>(w0.x, xs, neg, rc, &acc[0]);
                        word8<256 * jj + 16>(w0.y, xs, neg, rc, &acc[2]);
                        word8<256 * jj + 128>(w1.x, xs, neg, rc, &acc[0]);
                        word8<256 * jj + 144>(w1.y, xs, neg, rc, &acc[2]);
                    }
                }
            };
            if (nblk == kBPS) {                                      // full stage: no guards, the accumulators stay in place
                block_step(std::integral_constant<int, 0>{}); block_step(std::integral_constant<int, 1>{});
                block_step(std::integral_constant<int, 2>{}); block_step(std::integral_constant<int, 3>{});
                block_step(std::integral_constant<int, 4>{}); block_step(std::integral_constant<int, 5>{});
                block_step(std::integral_constant<int, 6>{}); block_step(std::integral_constant<int, 7>{});
            } else {
                if (nblk > 0) block_step(std::integral_constant<int, 0>{});
                if (nblk > 1) block_step(std::integral_constant<int, 1>{});
                if (nblk > 2) block_step(std::integral_constant<int, 2>{});
                if (nblk > 3) block_step(std::integral_constant<int, 3>{});
                if (nblk > 4) block_step(std::integral_constant<int, 4>{});
                if (nblk > 5) block_step(std::integral_constant<int, 5>{});
                if (nblk > 6) block_step(std::integral_constant<int, 6>{});
            }
            __syncwarp();
            if (lane == 0) {
                if (MBITS == 8) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp WROTE the stage (t = x * s) before the async-proxy refill
                issue();
            }
            s ^= 1u;
            if (s == 0) phase ^= 1u;
        }
        // the four accumulators of a row live in four threads: (acc_1 + acc_2) and (acc_3 + acc_4) across the pair that differs
        // in the low bit of k, their sum across the other pair, then the hadd tree (CloverBase.h:149-157); fp32 add commutes
        float tt[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) { tt[2 * p] = lo_f(acc[p]); tt[2 * p + 1] = hi_f(acc[p]); }
#pragma unroll
        for (int l = 0; l < 8; ++l) tt[l] = __fadd_rn(tt[l], __shfl_xor_sync(0xFFFFFFFFu, tt[l], 1));
#pragma unroll
        for (int l = 0; l < 8; ++l) tt[l] = __fadd_rn(tt[l], __shfl_xor_sync(0xFFFFFFFFu, tt[l], MBITS == 4 ? 2 : 16));
        if (k == 0)
            y[(uint64_t)item * kR8Rows + r] = __fadd_rn(__fadd_rn(__fadd_rn(tt[4], tt[0]), __fadd_rn(tt[6], tt[2])),
                                                        __fadd_rn(__fadd_rn(tt[5], tt[1]), __fadd_rn(tt[7], tt[3])));
    }
}

// One warp per row, lane 8k+l = the reference's chain (accumulator k, AVX lane l); plain loads. Same bits as the ring kernel.
template <int MBITS>
__global__ void __launch_bounds__(256)
k_mvm_f32_simple(const uint8_t *__restrict__ values, const float *__restrict__ scales, uint64_t rows, uint64_t cols,
                 const float *__restrict__ x, float *__restrict__ y) {
    constexpr float kQ = MBITS == 4 ? 7.0f : 127.0f;
    const int lane = threadIdx.x & 31, k = lane >> 3, l = lane & 7;
    const uint64_t hb = cols >> 6;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < rows; r += nwarps) {
        const uint8_t *row = values + (MBITS == 4 ? (r * cols) >> 1 : r * cols);
        const float *su = scales + (r >> 6) * hb;
        float acc = 0.f;
        for (uint64_t b = 0; b < hb; ++b) {
            const float s = __fdiv_rn(__ldg(su + b), kQ);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 32 * h + 8 * k + l;
                int q;
                if (MBITS == 4) {
                    const int byte = (int)(int8_t)row[b * 32 + (e >> 1)];
                    q = (e & 1) ? ((byte << 28) >> 28) : (byte >> 4);
                } else {
                    q = (int)(int8_t)row[b * 64 + e];
                }
                const float xv = __ldg(x + b * 64 + e), qf = __int2float_rn(q);
                if (MBITS == 4) acc = __fmaf_rn(xv, __fmul_rn(qf, s), acc);
                else            acc = __fmaf_rn(__fmul_rn(xv, s), qf, acc);
            }
        }
        acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, 8));             // acc_1+acc_2 | acc_3+acc_4
        acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, 16));            // sum_1 + sum_2
        acc = hadd8_butterfly(acc);
        if (lane == 0) y[r] = acc;
    }
}

template <int MBITS>
static int launch_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols, const float *x32,
                          float *y32, cudaStream_t stream) {
    if (rows == 0 || cols == 0) return CLOVER_OK;
    const char *impl = getenv("CLOVER_GEMV_IMPL");
    const bool simple = (impl && !strcmp(impl, "simple")) || (reinterpret_cast<uintptr_t>(values) & 15u) != 0 ||
                        (reinterpret_cast<uintptr_t>(x32) & 15u) != 0;
    if (simple) {
        const uint64_t want = (rows + 7) / 8, cap = (uint64_t)sm_count() * 8;
        k_mvm_f32_simple<MBITS><<<(unsigned)(want > cap ? cap : want), 256, 0, stream>>>(
            reinterpret_cast<const uint8_t *>(values), scales, rows, cols, x32, y32);
    } else {
        // 8-row items (k_mvm_f32_ring8, 28 warps per SM) unless the matrix is tall enough to fill the 16-row kernel's warp slots
        // (8-bit) / to fill them twice over (4-bit). Measured on B200 (tools/f32_sweep.py, us 16-row / 8-row items): 4-bit
        // 32768^2 198.7 / 182.6, 8192 x 32768 86.3 / 72.0, 32768 x 8192 55.0 / 51.6, 65536 x 16384 181.7 / 182.8; 8-bit 32768^2
        // 221.2 / 216.8, 8192 x 32768 131.1 / 100.8, 65536 x 16384 203.8 / 219.4. CLOVER_GEMV_IMPL=rows16 / rows8 force either.
        const bool rows8 = impl && !strcmp(impl, "rows8") ? true : impl && !strcmp(impl, "rows16") ? false
                                                          : rows / 16 <= (uint64_t)sm_count() * kF32Warps * (MBITS == 4 ? 2 : 1);
        CUtensorMap tmap;
        int rc = make_tensor_map_u8_2d_sw128(&tmap, values, rows, MBITS == 4 ? cols >> 1 : cols, rows8 ? kR8Rows : kF32Rows);
        if (rc != CLOVER_OK) return rc;
        // one CTA per SM as soon as there are that many work items: the warp-major item order then spreads the busy warps
        // evenly over all SMs (round 2a launched ceil(items / warps) CTAs and left 20 SMs idle at 32768 rows)
        const uint64_t nitems = rows / (rows8 ? kR8Rows : kF32Rows);
        const unsigned grid = (unsigned)(nitems < (uint64_t)sm_count() ? nitems : (uint64_t)sm_count());
        if (rows8) {
            auto kern = k_mvm_f32_ring8<MBITS>;
            CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kR8SmemBytes));   // per device and call: cheap
            kern<<<grid, kR8Warps * 32, kR8SmemBytes, stream>>>(tmap, scales, rows, cols, x32, y32, 0u);
        } else {
            const int smem = kF32SmemBytes;
            auto kern = k_mvm_f32_ring<MBITS>;
            CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            kern<<<grid, kF32Warps * 32, smem, stream>>>(tmap, scales, rows, cols, x32, y32, 0u);
        }
    }
    count_launch();
    return launch_status("k_mvm_f32");
}

}  // namespace clover

using namespace clover;

extern "C" {

int clover_m4_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      const float *x32, float *y32, void *stream) {
    CLOVER_REQUIRE(values && scales && x32 && y32, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID, "rows and cols must be multiples of 128 (include/CloverMatrix.h:48-50)");
    return launch_mvm_f32<4>(values, scales, rows, cols, x32, y32, (cudaStream_t)stream);
}

int clover_m8_mvm_f32(const int8_t *values, const float *scales, uint64_t rows, uint64_t cols,
                      const float *x32, float *y32, void *stream) {
    CLOVER_REQUIRE(values && scales && x32 && y32, CLOVER_ERR_INVALID, "null pointer");
    CLOVER_REQUIRE(rows % 128u == 0 && cols % 128u == 0, CLOVER_ERR_INVALID, "rows and cols must be multiples of 128 (include/CloverMatrix.h:48-50)");
    return launch_mvm_f32<8>(values, scales, rows, cols, x32, y32, (cudaStream_t)stream);
}

}  // extern "C"
