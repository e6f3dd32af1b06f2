// gemm4_tc3.cu - 4-bit GEMM, CTA-pair kernel with 256x192 tiles and FULL register staging of every K-slab.
//
// Same arithmetic as gemm4_tc.cu (E4M3-expanded operands, exact fp32 slab sums out of tcgen05.mma kind::f8f6f4,
// one fp32 FMA per output element per K-slab of 64, sequential in kb => bit-identical to k_gemm4_simt).
// What changes is the shape of the pipeline, following the two bounds measured in round 1
// (profiles/r01_gemm4_notes.md, profiles/r01_mma_probe.txt):
//
//   * operand delivery: a single CTA with a 128x256 tile needs 96 B/clk of operands per SM at full MMA rate and a
//     128x128 tile 128 B/clk; the TMA+MMA pipeline alone saturated at ~68 B/clk/SM. A CTA PAIR (cta_group::2)
//     computing a 256x192 tile needs 128 rows of A' + 96 rows of B' per CTA per slab = 14 KiB per 192 MMA cycles.
//   * TMEM hand-off: with 128x256 per CTA the tile's fp32 accumulators take half the register file, a thread can
//     land only half of its share of a slab, and a TMEM buffer goes back to the MMA warp only after half of the
//     slab's FMAs. With 128x192 per CTA (12 epilogue warps x 64 columns, 160 registers per thread after
//     setmaxnreg) every thread holds 64 accumulators AND two 32-column landing buffers: a slab leaves TMEM as soon
//     as its two tcgen05.ld have landed, before any of its FMAs, and the load of the next slab's first half
//     overlaps the FMAs of this slab's second half.
//
//   warp 0      TMA producer (both CTAs): own 128 rows of A', own 96 rows of B' per 128-byte K block, 7-stage ring
//   warp 1      MMA issuer (leader CTA): per stage 2 slabs x 2 tcgen05.mma.cta_group::2 (M256 N192 K32) into one of
//               two TMEM buffers (columns 0.. / 256..), multicast commits to both CTAs
//   warp 2      TMEM allocator
//   warps 4-15  epilogue: warp = 32 rows (lane quadrant) x 64 columns (one scale tile), thread = one row
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "runtime.cuh"
#include "tcgen05.cuh"

namespace clover {

constexpr int k3BM = 128;                         // rows per CTA (pair tile: 256)
constexpr int k3BN = 192;                         // columns of the pair tile; each CTA loads k3BN/2 rows of B'
constexpr int k3BK = 128;
constexpr int k3EpiWarps = 12;                    // 4 lane quadrants x 3 column blocks of 64
constexpr int k3Threads = 128 + 32 * k3EpiWarps;  // 512: launched with 128 registers/thread = 128*32 + 384*160 after setmaxnreg
constexpr int k3AStage = k3BM * k3BK, k3BStage = (k3BN / 2) * k3BK, k3StageBytes = k3AStage + k3BStage;   // 16 + 12 KiB
constexpr int k3Stages = 7;
constexpr int k3Smem = k3Stages * k3StageBytes + 1024 + 256;
constexpr uint32_t k3GroupM = 8;

__device__ __forceinline__ void tile_coords3(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t &tm, uint32_t &tn) {
    const uint32_t group_sz = k3GroupM * tiles_n;
    const uint32_t g = t / group_sz, r = t % group_sz;
    const uint32_t first = g * k3GroupM;
    const uint32_t gm = min(k3GroupM, tiles_m - first);
    tm = first + r % gm;
    tn = r / gm;
}

__device__ __forceinline__ void tma3_load_2d_2sm(uint32_t smem_dst, const void *tensor_map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(tensor_map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void umma3_commit_mc2(uint32_t bar) {     // arrive on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar3_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster3_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(k3Threads, 1)
k_gemm4_tc3(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const float *__restrict__ as, const float *__restrict__ bs, uint32_t M, uint32_t N, uint32_t K,
            float *__restrict__ c, uint64_t ldc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem + k3Stages * k3StageBytes;
    const uint32_t full = bars, empty = full + 8 * k3Stages, tfull = empty + 8 * k3Stages, tempty = tfull + 16;
    const uint32_t slot = tempty + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t tiles_m = (M + 2 * k3BM - 1) / (2 * k3BM), tiles_n = (N + k3BN - 1) / k3BN, ntiles = tiles_m * tiles_n;
    const uint32_t kblocks = K / k3BK, KB = K >> 6, NB = N >> 6;

    if (threadIdx.x == 0) {
        for (int i = 0; i < k3Stages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * i), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * i), "r"(1) : "memory");
        }
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tfull + 8 * b), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tempty + 8 * b), "r"(2 * k3EpiWarps) : "memory");
        }
        mbar_fence_init();
        tma_prefetch_descriptor(&map_a);
        tma_prefetch_descriptor(&map_b);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        tmem_relinquish<2>();
    }
    tc_fence_before();
    cluster3_sync_all();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));

    if (warp < 4) {
        reg_dealloc<32>();
        if (warp == 0) {
            // ===== TMA producer (both CTAs) =====
            const uint32_t full_leader = full & 0xFEFFFFFFu;           // same offset in the even CTA of the pair
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = pair; t < ntiles; t += npairs) {
                uint32_t tm, tn;
                tile_coords3(t, tiles_m, tiles_n, tm, tn);
                const int row_a = (int)(tm * 2 * k3BM + rank * k3BM), row_b = (int)(tn * k3BN + rank * (k3BN / 2));
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    mbar_wait_a(empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx_a(full + 8 * stage, 2 * k3StageBytes);
                        const uint32_t sa = smem + stage * k3StageBytes;
                        tma3_load_2d_2sm(sa, &map_a, (int)(kb * k3BK), row_a, full_leader + 8 * stage);
                        tma3_load_2d_2sm(sa + k3AStage, &map_b, (int)(kb * k3BK), row_b, full_leader + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == k3Stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1 && rank == 0) {
            // ===== MMA issuer (leader CTA): slab g -> TMEM buffer g & 1 =====
            const uint32_t idesc = umma_idesc(UMMA_E4M3, 2 * k3BM, k3BN);
            uint32_t stage = 0, phase = 0, g = 0;
            for (uint32_t t = pair; t < ntiles; t += npairs) {
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    mbar_wait_a(full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem + stage * k3StageBytes;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + k3AStage);
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h, ++g) {
                        const uint32_t b = g & 1;
                        mbar_wait_a(tempty + 8 * b, ((g >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem + b * 256;
                        if (elect_one()) {
                            umma_ss<UMMA_E4M3, 2>(d, da + 4 * h, db + 4 * h, idesc, 0);
                            umma_ss<UMMA_E4M3, 2>(d, da + 4 * h + 2, db + 4 * h + 2, idesc, 1);
                            umma3_commit_mc2(tfull + 8 * b);
                            if (h == 1) umma3_commit_mc2(empty + 8 * stage);     // smem stage free once its MMAs retire
                        }
                        __syncwarp();
                    }
                    if (++stage == k3Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: warps 4..15; warp = lane quadrant q x column block cb (64 columns = one scale tile) =====
        reg_alloc<160>();
        const uint32_t q = warp & 3, cb = (uint32_t)(warp - 4) >> 2;
        const uint32_t taddr = tmem + ((q * 32) << 16) + cb * 64;
        uint32_t tempty_leader;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(tempty_leader) : "r"(tempty), "r"(0));
        uint64_t acc[32];
        uint32_t ra[32], rb[32];
        uint32_t g = 0;                                                  // slab counter: buffer g & 1, parity (g >> 1) & 1
        for (uint32_t t = pair; t < ntiles; t += npairs) {
            uint32_t tm, tn;
            tile_coords3(t, tiles_m, tiles_n, tm, tn);
            const uint32_t row0 = tm * 2 * k3BM + rank * k3BM;          // first row of this CTA's half of the tile
            const uint32_t jb = tn * 3 + cb;                             // 64-column block of C
            const bool live = row0 < M && jb < NB;                       // M, N multiples of 128 resp. 64: all in or all out
            const float *pa = as + (uint64_t)(row0 < M ? (row0 >> 6) + (q >> 1) : 0) * KB;
            const float *pb = bs + (uint64_t)(jb < NB ? jb : 0) * KB;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0ull;
            float sv = 0.f;
            if (MODE == 0) {
                // ---- mode 0: rolling halves - one 32-column load in flight per warp ----
                {   // first half of the tile's first slab -> ra
                    const uint32_t b = g & 1;
                    mbar_wait_a(tfull + 8 * b, (g >> 1) & 1);
                    tc_fence_after();
                    tmem_ld32(taddr + 256 * b, ra);
                }
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    if ((kb & 31) == 0) {        // lane l owns the scale of slab kb + l: s = (sA * (1/49)) * sB, the reference's order
                        const uint32_t kl = min(kb + lane, KB - 1);
                        sv = __fmul_rn(__fmul_rn(__ldg(pa + kl), 1.0f / 49.0f), __ldg(pb + kl));
                    }
                    const float s = __shfl_sync(0xFFFFFFFFu, sv, kb & 31);
                    const uint32_t b = g & 1;
                    tmem_ld_wait(ra);                                        // first half of slab g has landed
                    tmem_ld32(taddr + 256 * b + 32, rb);                     // second half, lands under the FMAs below
#pragma unroll
                    for (int j = 0; j < 16; ++j) ffma2(acc[j], s, ra[2 * j], ra[2 * j + 1]);
                    tmem_ld_wait(rb);
                    tc_fence_before();                                       // this warp's share of the buffer is drained
                    if (lane == 0) mbar3_arrive_cluster(tempty_leader + 8 * b);
                    ++g;
                    if (kb + 1 < KB) {                                       // first half of the next slab, lands under the FMAs below
                        const uint32_t b2 = g & 1;
                        mbar_wait_a(tfull + 8 * b2, (g >> 1) & 1);
                        tc_fence_after();
                        tmem_ld32(taddr + 256 * b2, ra);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) ffma2(acc[16 + j], s, rb[2 * j], rb[2 * j + 1]);
                }
            } else if (MODE == 1) {
                // ---- mode 1: batch - both halves of a slab are requested together, the buffer is released before any FMA ----
                for (uint32_t kb = 0; kb < KB; ++kb, ++g) {
                    if ((kb & 31) == 0) {
                        const uint32_t kl = min(kb + lane, KB - 1);
                        sv = __fmul_rn(__fmul_rn(__ldg(pa + kl), 1.0f / 49.0f), __ldg(pb + kl));
                    }
                    const float s = __shfl_sync(0xFFFFFFFFu, sv, kb & 31);
                    const uint32_t b = g & 1;
                    mbar_wait_a(tfull + 8 * b, (g >> 1) & 1);
                    tc_fence_after();
                    tmem_ld32(taddr + 256 * b, ra);
                    tmem_ld32(taddr + 256 * b + 32, rb);
                    tmem_ld_wait(ra);
                    tmem_ld_wait(rb);
                    tc_fence_before();
                    if (lane == 0) mbar3_arrive_cluster(tempty_leader + 8 * b);
#pragma unroll
                    for (int j = 0; j < 16; ++j) ffma2(acc[j], s, ra[2 * j], ra[2 * j + 1]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) ffma2(acc[16 + j], s, rb[2 * j], rb[2 * j + 1]);
                }
            } else if (MODE == 3 || MODE == 4) {
                // ---- measurement only (wrong results): 3 = hand-off without loads or FMAs, 4 = batch loads without FMAs ----
                for (uint32_t kb = 0; kb < KB; ++kb, ++g) {
                    const uint32_t b = g & 1;
                    mbar_wait_a(tfull + 8 * b, (g >> 1) & 1);
                    tc_fence_after();
                    if (MODE == 4) {
                        tmem_ld32(taddr + 256 * b, ra);
                        tmem_ld32(taddr + 256 * b + 32, rb);
                        tmem_ld_wait(ra);
                        tmem_ld_wait(rb);
                        acc[kb & 31] ^= ((uint64_t)ra[kb & 31] << 32) | rb[kb & 31];
                    }
                    tc_fence_before();
                    if (lane == 0) mbar3_arrive_cluster(tempty_leader + 8 * b);
                }
            } else {
                // ---- mode 2: staggered batch - ra(g+1) is requested under the FMAs of rb(g), rb(g+1) right after them;
                //      one wait covers both, so two loads per warp overlap and the buffer is still released before its FMAs ----
                {
                    const uint32_t b = g & 1;
                    mbar_wait_a(tfull + 8 * b, (g >> 1) & 1);
                    tc_fence_after();
                    tmem_ld32(taddr + 256 * b, ra);
                    tmem_ld32(taddr + 256 * b + 32, rb);
                }
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    if ((kb & 31) == 0) {
                        const uint32_t kl = min(kb + lane, KB - 1);
                        sv = __fmul_rn(__fmul_rn(__ldg(pa + kl), 1.0f / 49.0f), __ldg(pb + kl));
                    }
                    const float s = __shfl_sync(0xFFFFFFFFu, sv, kb & 31);
                    tmem_ld_wait(ra);
                    tmem_ld_wait(rb);
                    tc_fence_before();                                       // slab g is in registers
                    if (lane == 0) mbar3_arrive_cluster(tempty_leader + 8 * (g & 1));
                    ++g;
                    const bool more = kb + 1 < KB;
                    const uint32_t b2 = g & 1;
#pragma unroll
                    for (int j = 0; j < 16; ++j) ffma2(acc[j], s, ra[2 * j], ra[2 * j + 1]);
                    if (more) {
                        mbar_wait_a(tfull + 8 * b2, (g >> 1) & 1);
                        tc_fence_after();
                        tmem_ld32(taddr + 256 * b2, ra);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) ffma2(acc[16 + j], s, rb[2 * j], rb[2 * j + 1]);
                    if (more) tmem_ld32(taddr + 256 * b2 + 32, rb);
                }
            }
            if (live) {
                float *crow = c + (uint64_t)(row0 + q * 32 + lane) * ldc + (uint64_t)jb * 64;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float4 o;
                    o.x = __uint_as_float((uint32_t)acc[2 * j]);     o.y = __uint_as_float((uint32_t)(acc[2 * j] >> 32));
                    o.z = __uint_as_float((uint32_t)acc[2 * j + 1]); o.w = __uint_as_float((uint32_t)(acc[2 * j + 1] >> 32));
                    *reinterpret_cast<float4 *>(crow + 4 * j) = o;
                }
            }
        }
    }
    tc_fence_before();
    cluster3_sync_all();     // nobody leaves while the peer's MMAs / remote arrivals may still touch this CTA
    if (warp == 2) tmem_dealloc<2>(tmem, 512);
}

int gemm4_tc3_expanded(const uint8_t *a8, const float *as, const uint8_t *b8, const float *bs, uint64_t M, uint64_t N,
                       uint64_t K, float *c, uint64_t ldc, cudaStream_t stream) {
    CUtensorMap map_a, map_b;
    int rc = make_tensor_map_u8_2d_sw128(&map_a, a8, M, K, k3BM);
    if (rc != CLOVER_OK) return rc;
    rc = make_tensor_map_u8_2d_sw128(&map_b, b8, N, K, k3BN / 2);
    if (rc != CLOVER_OK) return rc;
    const int mode = [] { const char *e = getenv("CLOVER_GEMM_EPI"); return e ? atoi(e) : 1; }();   // epilogue schedule, see the kernel (1 = batch: best measured)
    auto kern = mode == 0 ? k_gemm4_tc3<0> : mode == 1 ? k_gemm4_tc3<1> : mode == 3 ? k_gemm4_tc3<3> : mode == 4 ? k_gemm4_tc3<4> : k_gemm4_tc3<2>;
    CLOVER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, k3Smem));
    const uint64_t ntiles = ((M + 2 * k3BM - 1) / (2 * k3BM)) * ((N + k3BN - 1) / k3BN);
    const unsigned groups = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sm_count() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * groups);
    cfg.blockDim = dim3(k3Threads);
    cfg.dynamicSmemBytes = k3Smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CLOVER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, as, bs, (uint32_t)M, (uint32_t)N, (uint32_t)K, c, ldc));
    count_launch();
    return launch_status("k_gemm4_tc3");
}

}  // namespace clover
