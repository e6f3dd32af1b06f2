// tcgen05.cuh - thin inline-PTX layer over the 5th-generation tensor core (tcgen05.mma / TMEM) for the
// 4-bit GEMM. One elected thread issues tcgen05.mma; operands are K-major shared-memory tiles written by
// TMA with the 128-byte swizzle; accumulators live in tensor memory and are drained with tcgen05.ld.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables and were
// validated on a B200 by tools/mma_probe.cu (kind::i8 and kind::f8f6f4 both exact on the Clover domain).
#pragma once
#include <stdint.h>
#include "async_copy.cuh"

namespace clover {

template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    else         asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
template <int CG> __device__ __forceinline__ void tmem_relinquish() {
    if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
    else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile whose rows are 128 bytes, laid out by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B: 8-row atoms of 1024 B stacked along M/N (stride byte offset 1024).
// Advancing K by 32 bytes inside the 128-byte row = +2 on the (16-byte granular) start-address field.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset      bits [16,30) (unused here)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B bits [61,64)
    return d;
}

enum UmmaKind { UMMA_I8 = 0, UMMA_E4M3 = 1 };
// Instruction descriptor: c_format [4,6), a_format [7,10), b_format [10,13), both operands K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int kind, int M, int N) {
    return (kind == UMMA_I8 ? (2u << 4) | (1u << 7) | (1u << 10)      // s32 += s8 * s8
                            : (1u << 4) | (0u << 7) | (0u << 10)) |   // f32 += e4m3 * e4m3
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int KIND, int CG>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if (KIND == UMMA_I8 && CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    if (KIND == UMMA_E4M3 && CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    if (KIND == UMMA_I8 && CG == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    if (KIND == UMMA_E4M3 && CG == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
template <int CG> __device__ __forceinline__ void umma_commit(uint64_t *bar) {
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else         asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CG> __device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else         asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of 32-bit accumulators: thread t of the warp receives lane (base_lane + t),
// columns col .. col+31. The registers are NOT valid until tmem_ld_wait() on the same array.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
// tcgen05.wait::ld with the destination registers threaded through as in/out operands, so that no use of
// r[] can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

// acc.xy = s * d.xy + acc.xy, both halves rounded like a scalar fma.rn (SASS: FFMA2 with a scalar-broadcast operand)
__device__ __forceinline__ void ffma2(uint64_t &acc, float s, uint32_t d_lo, uint32_t d_hi) {
    const uint64_t d = ((uint64_t)d_hi << 32) | d_lo;
    const uint64_t ss = ((uint64_t)__float_as_uint(s) << 32) | __float_as_uint(s);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ss), "l"(d));
}

// One lane of a fully converged warp. The producer / MMA warps run their loops warp-uniformly and only the
// issue of the TMA / tcgen05 instruction is predicated on this: under `if (lane == 0)` control flow ptxas
// cannot prove descriptors uniform and wraps every UTCQMMA / UTMALDG in an ELECT + R2UR "waterfall" loop
// (~15 dependent instructions per MMA), which throttles the single issuing thread.
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}

template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace clover
