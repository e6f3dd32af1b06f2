// tools/mma_probe.cu - measurement probe for the tcgen05 GEMM design (not part of the product library).
//
//   mma_probe check   one CTA: TMA (SWIZZLE_128B) -> smem -> tcgen05.mma kind::i8 / kind::f8f6f4(E4M3) -> TMEM ->
//                     tcgen05.ld -> global; compared with the exact integer product on the host. Validates the
//                     shared-memory / instruction descriptors and whether E4M3 x E4M3 -> fp32 accumulation is
//                     EXACT for the Clover domain (|q| <= 7, K-slab sums <= 64*49).
//   mma_probe peak    bare issue loop, all SMs: measured tensor-pipe peak for kind::i8 and kind::f8f6f4,
//                     cta_group::1 (M128 N256 K32) and cta_group::2 (M256 N256 K32); burst and sustained.
//   mma_probe epi     epilogue ceiling: 8 warps draining 128x256 fp32 from TMEM and applying fma.rn.f32x2.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/mma_probe tools/mma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) { } }
__device__ __forceinline__ void tma_load_2d(void *dst, const void *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B stacked along M/N (SBO = 1024 B)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;              // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;    // SBO
    d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}
// KIND: 0 = i8 (s8 x s8 -> s32), 1 = f8f6f4 (e4m3 x e4m3 -> f32)
__host__ __device__ constexpr uint32_t make_idesc(int kind, int M, int N) {
    return (kind == 0 ? (2u << 4) | (1u << 7) | (1u << 10) : (1u << 4) | (0u << 7) | (0u << 10)) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND, int CG>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (KIND == 0 && CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 1 && CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 0 && CG == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    if (KIND == 1 && CG == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int CG> __device__ __forceinline__ void mma_commit(uint64_t *bar) {
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else         asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// =========================================================================================================
// check: D[128 x 256] = A[128 x 128] * B[256 x 128]^T, one CTA, 128 threads
// =========================================================================================================
template <int KIND>
__global__ void __launch_bounds__(128) k_check(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                               uint32_t *out /* [128][256] raw 32-bit */) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 128 * 128;
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc<1>(&tmem_slot, 256); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_full, 128 * 128 + 256 * 128);
        tma_load_2d(sa, &map_a, 0, 0, &bar_full);
        tma_load_2d(sb, &map_b, 0, 0, &bar_full);
        mbar_wait(&bar_full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(KIND, 128, 256);
        const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss<KIND, 1>(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
        mma_commit<1>(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < 256; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[row * 256 + c + j] = r[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 256);
}

// =========================================================================================================
// check2: cta_group::2. D[256 x 128] = A[256 x 128] * B[128 x 128]^T on a CTA pair: CTA r holds A rows
// [128r, 128r+128) and B rows [64r, 64r+64); both CTAs' TMA loads signal the LEADER's barrier
// (.cta_group::2, peer bit cleared); the leader issues the MMAs and commits to both CTAs (multicast).
// =========================================================================================================
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const void *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}
template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k_check2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                uint32_t *out /* [256][128] */) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 128 * 128;
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc<2>(&tmem_slot, 128); tmem_relinquish<2>(); }
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        if (rank == 0) mbar_arrive_expect_tx(&bar_full, 2 * (128 * 128 + 64 * 128));
        tma_load_2d_2sm(sa, &map_a, 0, (int)rank * 128, &bar_full);
        tma_load_2d_2sm(sb, &map_b, 0, (int)rank * 64, &bar_full);
        if (rank == 0) {
            mbar_wait(&bar_full, 0);
            tc_fence_after();
            const uint32_t idesc = make_idesc(KIND, 256, 128);
            const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss<KIND, 2>(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar_mma)), "h"((uint16_t)3) : "memory");
        }
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = (int)rank * 128 + warp * 32 + lane;
    for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[row * 128 + c + j] = r[j];
    }
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) tmem_dealloc<2>(tmem, 128);
}

// =========================================================================================================
// peak: every CTA (or CTA pair) issues `iters` x 4 MMAs on resident shared memory
// =========================================================================================================
template <int KIND, int CG, int N>
__global__ void __launch_bounds__(128) k_peak(int iters, uint32_t seed, uint32_t *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // A: 128 rows x 128 B; B: N/CG rows x 128 B   (per CTA)
    uint8_t *sa = smem, *sb = smem + 128 * 128;
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    uint32_t cta_rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    // fill with Clover-domain data: int8 in [-7,7]*16 for i8, E4M3 codes of -7..7 for f8
    {
        uint32_t s = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
        const uint8_t e4m3[8] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E};
        for (int i = threadIdx.x; i < (128 + N / CG) * 128; i += blockDim.x) {
            s = s * 1664525u + 1013904223u;
            const int q = (int)((s >> 16) % 15u) - 7;
            smem[i] = KIND == 0 ? (uint8_t)(int8_t)(q * 16) : (uint8_t)(e4m3[q < 0 ? -q : q] | (q < 0 ? 0x80 : 0));
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc<CG>(&tmem_slot, 512); tmem_relinquish<CG>(); }
    tc_fence_before();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0 && cta_rank == 0) {
        const uint32_t idesc = make_idesc(KIND, 128 * CG, N);
        const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (it & 1) * 256;
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ss<KIND, CG>(d, da + 2 * k, db + 2 * k, idesc, (k & 1));   // 2 slabs of K=64
        }
        mma_commit<CG>(&bar_mma);
        mbar_wait(&bar_mma, 0);
        tc_fence_after();
    }
    tc_fence_before();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    if (warp == 0) tmem_dealloc<CG>(tmem, 512);
    if (sink && threadIdx.x == 0 && iters < 0) sink[blockIdx.x] = tmem;
}

// =========================================================================================================
// epi: the epilogue's ceiling. 8 warps; per "slab" each thread loads 128 fp32 columns of its TMEM lane
// (4 x tcgen05.ld.32x32b.x32) and applies acc = fma(s, d, acc) with packed fma.rn.f32x2 (PACKED=1) or scalar FFMA.
// =========================================================================================================
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ void fma2(uint64_t &acc, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

template <int PACKED>
__global__ void __launch_bounds__(256, 1) k_epi(int slabs, float s0, float *out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc<1>(&tmem_slot, 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col_base = (warp >> 2) * 128;
    uint64_t acc2[64];
    float acc[128];
    if (PACKED) { for (int i = 0; i < 64; ++i) acc2[i] = 0; } else { for (int i = 0; i < 128; ++i) acc[i] = 0.f; }
    float s = s0;
    for (int sl = 0; sl < slabs; ++sl) {
        const uint32_t buf = (sl & 1) * 256;
        const uint64_t s2 = pack2(s, s);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem + lane_base + buf + col_base + c * 32, r);
            tc_wait_ld();
            if (PACKED) {
#pragma unroll
                for (int j = 0; j < 16; ++j) fma2(acc2[c * 16 + j], s2, ((uint64_t)r[2 * j + 1] << 32) | r[2 * j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c * 32 + j] = __fmaf_rn(s, __uint_as_float(r[j]), acc[c * 32 + j]);
            }
        }
        s += 1e-7f;
    }
    float t = 0.f;
    if (PACKED) { for (int i = 0; i < 64; ++i) t += __uint_as_float((uint32_t)acc2[i]) + __uint_as_float((uint32_t)(acc2[i] >> 32)); }
    else { for (int i = 0; i < 128; ++i) t += acc[i]; }
    out[blockIdx.x * 256 + threadIdx.x] = t;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 512);
    (void)lane;
}

// =========================================================================================================
// lat: hand-off latency of the MMA -> epilogue -> MMA loop through ONE TMEM slot (depth-1 ring), one CTA.
//   warp 1 (one elected lane): wait tempty; NMMA x tcgen05.mma (M128 x N, K32); tcgen05.commit -> tfull
//   warps 4..7: wait tfull; [LDTM x32 + wait::ld]; arrive tempty (count 4)
// cycles/iteration - NMMA * (N/2) = commit + wake-up + TMEM load + arrive + wake-up latencies.
// =========================================================================================================
template <int N, int NMMA, int WITH_LD>
__global__ void __launch_bounds__(256) k_lat(int iters, long long *out_cycles, uint32_t *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 128 * 128;
    __shared__ uint64_t tfull, tempty;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + N) * 128; i += blockDim.x) smem[i] = 0x38;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(&tfull, 1); mbar_init(&tempty, 4); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc<1>(&tmem_slot, 256); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t acc = 0;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, 128, N);
        const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&tempty, (it & 1) ^ 1);
            tc_fence_after();
            uint32_t el;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
            if (el) {
#pragma unroll
                for (int k = 0; k < NMMA; ++k) mma_ss<1, 1>(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, k > 0);
                mma_commit<1>(&tfull);
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
    } else if (warp >= 4) {
        long long t_ld = 0, t_wait = 0;
        for (int it = 0; it < iters; ++it) {
            const long long ta = clock64();
            mbar_wait(&tfull, it & 1);
            tc_fence_after();
            const long long tb = clock64();
            t_wait += tb - ta;
            if (WITH_LD) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (WITH_LD == 2 ? 192 : 0), r);   // 2: columns the MMA never writes
                tc_wait_ld();
                acc += r[0] ^ r[31];
            }
            tc_fence_before();
            t_ld += clock64() - tb;
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty)) : "memory");
            __syncwarp();
        }
        if (warp == 4 && lane == 0) { out_cycles[148 + blockIdx.x] = t_ld; out_cycles[296 + blockIdx.x] = t_wait; }
        if (sink && acc == 0x12345678u) sink[threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 256);
}

template <int N, int NMMA, int WITH_LD> static void run_lat(const char *name) {
    const int iters = 20000;
    long long *dcyc; CK(cudaMalloc(&dcyc, 3 * 148 * 8));
    const int smem = (128 + N) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_lat<N, NMMA, WITH_LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int grid : {1, 148}) {
        k_lat<N, NMMA, WITH_LD><<<grid, 256, smem>>>(iters, dcyc, nullptr);
        CK(cudaDeviceSynchronize());
        long long h[3 * 148];
        CK(cudaMemcpy(h, dcyc, 3 * 148 * 8, cudaMemcpyDeviceToHost));
        double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i] / iters; mean /= grid;
        printf("lat %-34s grid=%3d: %.1f cycles/iteration (MMA work %d cycles -> hand-off overhead %.1f); epilogue warp: %.1f in ld+wait::ld+fence, %.1f waiting for tfull\n",
               name, grid, mean, NMMA * N / 2, mean - NMMA * N / 2, (double)h[148] / iters, (double)h[296] / iters);
    }
    cudaFree(dcyc);
}

// =========================================================================================================
// ldtm: TMEM -> register bandwidth. NW warps, each looping over `reps` batches of 4 x tcgen05.ld.32x32b.x32
// (its 32 lanes x 128 columns = 16 KiB) followed by one tcgen05.wait::ld.
// =========================================================================================================
template <int NW, int V = 0, int NLD = 4>
__global__ void __launch_bounds__(NW * 32) k_ldtm(int reps, long long *out_cycles, uint32_t *sink) {
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t dummy_bar;
    if (threadIdx.x == 0) { mbar_init(&dummy_bar, (1u << 20) - 1); mbar_fence_init(); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc<1>(&tmem_slot, 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        uint32_t a[32], b[32], c[32], d[32];
        if (V & 2) tc_fence_after();
        tmem_ld32(base, a);
        if (NLD == 4) { tmem_ld32(base + 32, b); tmem_ld32(base + 64, c); tmem_ld32(base + 96, d); }
        tc_wait_ld();
        if (V & 1) tc_fence_before();
        if ((V & 4) && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&dummy_bar)) : "memory");
        acc += NLD == 4 ? (a[0] ^ b[1] ^ c[2] ^ d[3]) : (a[0] ^ a[31]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    if (sink && acc == 0x12345678u) sink[threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 512);
}
template <int NW, int V, int NLD> static void run_ldtm_v(const char *name) {
    const int reps = 20000;
    long long *dcyc; CK(cudaMalloc(&dcyc, 148 * 8));
    k_ldtm<NW, V, NLD><<<148, NW * 32>>>(reps, dcyc, nullptr);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, dcyc, 148 * 8, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < 148; ++i) mean += (double)h[i] / reps; mean /= 148;
    printf("ldtm-loop %2d warps, %d x ld.x32 + wait::ld %-44s: %.1f cycles per iteration\n", NW, NLD, name, mean);
    cudaFree(dcyc);
}
// ldtm + mma: NW load warps hammer TMEM columns 256..511 while one extra warp issues back-to-back MMAs
// (M128 N128 K32, e4m3) into columns 0..255. Reports both rates: do accumulator traffic and tcgen05.ld interfere?
template <int NW>
__global__ void __launch_bounds__(NW * 32 + 32) k_ldtm_mma(int reps, int mma_iters, long long *out_cycles, uint32_t *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 128; i += blockDim.x) smem[i] = 0x38 + (i & 7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc<1>(&tmem_slot, 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp == NW) {
        if (threadIdx.x == NW * 32) {
            const uint32_t idesc = make_idesc(1, 128, 128);
            const uint64_t da = desc_sw128(smem_u32(smem)), db = desc_sw128(smem_u32(smem + 128 * 128));
            for (int it = 0; it < mma_iters; ++it) {
                const uint32_t d = tmem + (it & 1) * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_ss<1, 1>(d, da + 2 * k, db + 2 * k, idesc, (k & 1));
            }
            mma_commit<1>(&bar_mma);
            mbar_wait(&bar_mma, 0);
            out_cycles[blockIdx.x * 2 + 1] = clock64() - t0;
        }
    } else {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 1) * 128;
        for (int it = 0; it < reps; ++it) {
            uint32_t a[32], b[32], c[32], d[32];
            tmem_ld32(base, a); tmem_ld32(base + 32, b); tmem_ld32(base + 64, c); tmem_ld32(base + 96, d);
            tc_wait_ld();
            acc += a[0] ^ b[1] ^ c[2] ^ d[3];
        }
        if (threadIdx.x == 0) out_cycles[blockIdx.x * 2] = clock64() - t0;
    }
    if (sink && acc == 0x12345678u) sink[threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tmem, 512);
}
template <int NW> static void run_ldtm_mma() {
    const int reps = 20000, mma_iters = 20000;     // 80000 MMAs of 64 cycles = 5.1M cycles if unimpeded
    long long *dcyc; CK(cudaMalloc(&dcyc, 148 * 16));
    const int smem = 256 * 128 + 1024;
    CK(cudaFuncSetAttribute(k_ldtm_mma<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_ldtm_mma<NW><<<148, NW * 32 + 32, smem>>>(reps, mma_iters, dcyc, nullptr);
    CK(cudaDeviceSynchronize());
    long long h[296];
    CK(cudaMemcpy(h, dcyc, 148 * 16, cudaMemcpyDeviceToHost));
    double ld = 0, mm = 0; for (int i = 0; i < 148; ++i) { ld += (double)h[2 * i] / reps; mm += (double)h[2 * i + 1] / (4.0 * mma_iters); }
    ld /= 148; mm /= 148;
    printf("ldtm+mma %2d load warps: %.1f cycles per 16 KiB batch per warp (%.0f B/clk/SM) while MMAs (M128N128K32, ideal 64 cycles) take %.1f cycles each\n",
           NW, ld, (double)NW * 16384 / ld, mm);
    cudaFree(dcyc);
}

template <int NW> static void run_ldtm() {
    const int reps = 20000;
    long long *dcyc; CK(cudaMalloc(&dcyc, 148 * 8));
    k_ldtm<NW><<<148, NW * 32>>>(reps, dcyc, nullptr);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, dcyc, 148 * 8, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < 148; ++i) mean += (double)h[i] / reps; mean /= 148;
    const double bytes = (double)NW * 32 * 128 * 4;
    printf("ldtm %2d warps: %.1f cycles per batch (16 KiB per warp) -> %.0f B/clk/SM ; a 128x128 fp32 slab (64 KiB) needs %.0f cycles\n",
           NW, mean, bytes / mean, 65536.0 / (bytes / mean));
    cudaFree(dcyc);
}

// =========================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return (EncodeTiledFn)fn;
}
static CUtensorMap make_map(const void *base, uint64_t rows, uint64_t row_bytes, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {row_bytes, rows};
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {128, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
    return m;
}

static const uint8_t kE4M3[8] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E};

template <int KIND> static int run_check() {
    const int M = 128, N = 256, K = 128;
    std::vector<int> qa(M * K), qb(N * K);
    std::vector<uint8_t> ha(M * K), hb(N * K);
    uint32_t s = 12345u + KIND;
    auto next = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 16) % 15u) - 7; };
    for (int i = 0; i < M * K; ++i) { int q = next(); qa[i] = q; ha[i] = KIND == 0 ? (uint8_t)(int8_t)(16 * q) : (uint8_t)(kE4M3[abs(q)] | (q < 0 ? 0x80 : 0)); }
    for (int i = 0; i < N * K; ++i) { int q = next(); qb[i] = q; hb[i] = KIND == 0 ? (uint8_t)(int8_t)(16 * q) : (uint8_t)(kE4M3[abs(q)] | (q < 0 ? 0x80 : 0)); }
    uint8_t *da, *db; uint32_t *dout;
    CK(cudaMalloc(&da, M * K)); CK(cudaMalloc(&db, N * K)); CK(cudaMalloc(&dout, M * N * 4));
    CK(cudaMemcpy(da, ha.data(), M * K, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), N * K, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, M * N * 4));
    CUtensorMap ma = make_map(da, M, K, M), mb = make_map(db, N, K, N);
    const int smem = (128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_check<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_check<KIND><<<1, 128, smem>>>(ma, mb, dout);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> ho(M * N);
    CK(cudaMemcpy(ho.data(), dout, M * N * 4, cudaMemcpyDeviceToHost));
    long bad = 0; int shown = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            long ref = 0;
            for (int k = 0; k < K; ++k) ref += qa[i * K + k] * qb[j * K + k];
            double got;
            if (KIND == 0) got = (double)(int32_t)ho[i * N + j] / 256.0;
            else { float f; memcpy(&f, &ho[i * N + j], 4); got = f; }
            if (got != (double)ref) { ++bad; if (shown++ < 8) printf("  mismatch [%d][%d] got %.4f want %ld (raw %08x)\n", i, j, got, ref, ho[i * N + j]); }
        }
    printf("check kind=%s: %ld mismatches of %d (%s)\n", KIND == 0 ? "i8" : "f8f6f4/e4m3", bad, M * N, bad ? "FAIL" : "EXACT");
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return bad != 0;
}

template <int KIND> static int run_check2() {
    const int M = 256, N = 128, K = 128;
    std::vector<int> qa(M * K), qb(N * K);
    std::vector<uint8_t> ha(M * K), hb(N * K);
    uint32_t s = 777u + KIND;
    auto next = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 16) % 15u) - 7; };
    for (int i = 0; i < M * K; ++i) { int q = next(); qa[i] = q; ha[i] = KIND == 0 ? (uint8_t)(int8_t)(16 * q) : (uint8_t)(kE4M3[abs(q)] | (q < 0 ? 0x80 : 0)); }
    for (int i = 0; i < N * K; ++i) { int q = next(); qb[i] = q; hb[i] = KIND == 0 ? (uint8_t)(int8_t)(16 * q) : (uint8_t)(kE4M3[abs(q)] | (q < 0 ? 0x80 : 0)); }
    uint8_t *da, *db; uint32_t *dout;
    CK(cudaMalloc(&da, M * K)); CK(cudaMalloc(&db, N * K)); CK(cudaMalloc(&dout, M * N * 4));
    CK(cudaMemcpy(da, ha.data(), M * K, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), N * K, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, M * N * 4));
    CUtensorMap ma = make_map(da, M, K, 128), mb = make_map(db, N, K, 64);
    const int smem = (128 + 64) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_check2<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_check2<KIND><<<2, 128, smem>>>(ma, mb, dout);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> ho(M * N);
    CK(cudaMemcpy(ho.data(), dout, M * N * 4, cudaMemcpyDeviceToHost));
    long bad = 0; int shown = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            long ref = 0;
            for (int k = 0; k < K; ++k) ref += qa[i * K + k] * qb[j * K + k];
            double got;
            if (KIND == 0) got = (double)(int32_t)ho[i * N + j] / 256.0;
            else { float f; memcpy(&f, &ho[i * N + j], 4); got = f; }
            if (got != (double)ref) { ++bad; if (shown++ < 8) printf("  mismatch [%d][%d] got %.4f want %ld (raw %08x)\n", i, j, got, ref, ho[i * N + j]); }
        }
    printf("check2 (cta_group::2, M256 N128) kind=%s: %ld mismatches of %d (%s)\n", KIND == 0 ? "i8" : "f8f6f4/e4m3", bad, M * N, bad ? "FAIL" : "EXACT");
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return bad != 0;
}

template <int KIND, int CG, int N = 256> static void run_peak(const char *name, int iters, int reps) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem = (128 + N / CG) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_peak<KIND, CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / CG * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, k_peak<KIND, CG, N>, 64, 1u, (uint32_t *)nullptr));   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f, total = 0.f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, k_peak<KIND, CG, N>, iters, (uint32_t)(r + 2), (uint32_t *)nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best; total += ms;
    }
    const double ops = 2.0 * (128.0 * CG) * N * 32 * 4 * iters * (cfg.gridDim.x / CG);
    printf("peak %-28s iters=%d grid=%u: best %.3f ms -> %.1f TOPS ; mean over %d reps (%.0f ms total) -> %.1f TOPS\n", name, iters,
           cfg.gridDim.x, best, ops / best * 1e-9, reps, total, ops * reps / total * 1e-9);
}

template <int PACKED> static void run_epi(int slabs) {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    float *out; CK(cudaMalloc(&out, sms * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_epi<PACKED><<<sms, 256>>>(16, 1.0f, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        k_epi<PACKED><<<sms, 256>>>(slabs, 1.0f, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double ns_per_slab = best * 1e6 / slabs;
    printf("epi %s: %d slabs of 128x256 per SM in %.3f ms -> %.1f ns/slab (= %.0f cycles at %d MHz max clock); "
           "an i8/f8 slab of K=64 at 8192 MAC/clk/SM is 256 cycles\n", PACKED ? "fma.rn.f32x2" : "scalar FFMA", slabs, best,
           ns_per_slab, ns_per_slab * khz * 1e-6, khz / 1000);
    cudaFree(out);
}

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "check";
    if (!strcmp(mode, "check")) {
        int bad = run_check<0>();
        bad |= run_check<1>();
        bad |= run_check2<0>();
        bad |= run_check2<1>();
        return bad;
    }
    if (!strcmp(mode, "peak")) {
        const int burst = 20000, sustained = argc > 2 ? atoi(argv[2]) : 400000;
        run_peak<0, 1>("i8 cta_group::1 M128N256", burst, 5);
        run_peak<1, 1>("e4m3 cta_group::1 M128N256", burst, 5);
        run_peak<0, 2>("i8 cta_group::2 M256N256", burst, 5);
        run_peak<1, 2>("e4m3 cta_group::2 M256N256", burst, 5);
        run_peak<1, 1, 128>("e4m3 cta_group::1 M128N128", 2 * burst, 5);
        run_peak<1, 2, 128>("e4m3 cta_group::2 M256N128", 2 * burst, 5);
        run_peak<1, 1, 64>("e4m3 cta_group::1 M128N64", 4 * burst, 5);
        run_peak<0, 1, 128>("i8 cta_group::1 M128N128", 2 * burst, 5);
        if (sustained <= 0) return 0;
        run_peak<0, 2>("i8 cg2 sustained", sustained, 8);
        run_peak<1, 2>("e4m3 cg2 sustained", sustained, 8);
        return 0;
    }
    if (!strcmp(mode, "epi")) {
        run_epi<1>(20000);
        run_epi<0>(20000);
        return 0;
    }
    if (!strcmp(mode, "ldtmv")) {
        run_ldtm_v<4, 0, 1>("");
        run_ldtm_v<4, 1, 1>("+ fence::before_thread_sync");
        run_ldtm_v<4, 2, 1>("+ fence::after_thread_sync");
        run_ldtm_v<4, 3, 1>("+ both fences");
        run_ldtm_v<4, 7, 1>("+ both fences + mbarrier.arrive");
        run_ldtm_v<16, 7, 1>("+ both fences + mbarrier.arrive");
        run_ldtm_v<4, 7, 4>("+ both fences + mbarrier.arrive");
        return 0;
    }
    if (!strcmp(mode, "ldtm")) { run_ldtm<4>(); run_ldtm<8>(); run_ldtm<16>(); run_ldtm_mma<4>(); run_ldtm_mma<8>(); run_ldtm_mma<16>(); return 0; }
    if (!strcmp(mode, "lat")) {
        run_lat<128, 2, 1>("N128 x2 MMA, with tcgen05.ld");
        run_lat<128, 2, 0>("N128 x2 MMA, no tcgen05.ld");
        run_lat<128, 2, 2>("N128 x2 MMA, tcgen05.ld of OTHER columns");
        run_lat<256, 2, 1>("N256 x2 MMA, with tcgen05.ld");
        run_lat<32, 1, 0>("N32 x1 MMA, no tcgen05.ld");
        run_lat<32, 1, 1>("N32 x1 MMA, with tcgen05.ld");
        return 0;
    }
    printf("usage: mma_probe check|peak|epi|lat\n");
    return 1;
}
