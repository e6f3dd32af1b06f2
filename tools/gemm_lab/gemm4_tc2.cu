// gemm4_tc2.cu - EXPERIMENTAL variants of the 4-bit GEMM pipeline: a 4-deep ring of 128-column TMEM slots.
//
// Same arithmetic as gemm4_tc.cu (E4M3-expanded operands, exact fp32 slab sums, one fp32 FMA per output element
// per K-slab of 64, sequential in kb => bit-identical to k_gemm4_simt). Selected with CLOVER_GEMM_KERNEL:
//   pipe1  cta_group::1, tile 128x128, all barriers local
//   pair   cta_group::2, CTA pair shares a 256x128 tile (B' split across the pair, .multicast commits,
//          remote arrives on the leader's `tempty`)
// Both keep a slab in only 128 TMEM columns so that four slabs fit, give every epilogue thread two landing
// buffers and release a TMEM slot as soon as its tcgen05.ld has landed (before the FMAs).
// Measured on B200 at 16384^3 (profiles/r01_gemm4_notes.md): 1.41 POPS (pipe1) / 1.47 POPS (pair) against
// 2.03 POPS for the default kernel. Why they lose, from the in-kernel trace below (CLOVER_GEMM_DBG=4):
//   * 128-wide tiles need 128 (pipe1) / 96 (pair) bytes of operands per clock at full MMA rate; without any
//     epilogue the TMA+MMA pipeline alone tops out at 2.4 / 2.5 POPS (L2->SM + shared-memory bandwidth);
//   * every epilogue warp has to touch every slab, and one touch (mbarrier wait -> tcgen05.ld -> wait::ld ->
//     arrive) is a ~370-cycle dependent chain with 16 warps contending for TMEM, against 128 cycles of MMA
//     per slab; hiding it needs ~3 slabs of landing registers per thread, which the register file cannot
//     give next to the tile's accumulators.
// They stay in the build (tested against the DP4A kernel) as the starting point for round-2 work.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "runtime.cuh"
#include "tcgen05.cuh"

namespace clover {

constexpr int k2BM = 128;              // rows per CTA (CG = 2: the pair's tile has 256)
constexpr int k2BN = 128;              // columns of the tile; with CG = 2 each CTA loads k2BN/2 rows of B'
constexpr int k2BK = 128;
constexpr int k2EpiWarps = 16;
constexpr int k2Threads = 128 + 32 * k2EpiWarps;      // 640: launched with 96 registers/thread = 128*32 + 512*112 after setmaxnreg
constexpr uint32_t k2GroupM = 8;
// CG = cta_group: 2 = CTA pair (tile 256x128, B' split, cross-CTA barriers), 1 = single CTA (tile 128x128, local barriers)
template <int CG> struct PipeCfg {
    static constexpr int kAStage = k2BM * k2BK, kBStage = (k2BN / CG) * k2BK, kStageBytes = kAStage + kBStage;   // 24 / 32 KiB
    static constexpr int kStages = CG == 2 ? 8 : 6;                                                            // 192 KiB of operands
    static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ void tile_coords2(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t &tm, uint32_t &tn) {
    const uint32_t group_sz = k2GroupM * tiles_n;
    const uint32_t g = t / group_sz, r = t % group_sz;
    const uint32_t first = g * k2GroupM;
    const uint32_t gm = min(k2GroupM, tiles_m - first);
    tm = first + r % gm;
    tn = r / gm;
}

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const void *tensor_map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(tensor_map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar) {     // arrive on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");   // default .release.cta: a cluster-scope release would cost MEMBAR.GPU per slab, and no memory is handed over here (TMEM ordering is tcgen05.fence)
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CLOVER_GEMM_DBG=4: CTA 0 records clock64() of the hand-off events of its first kTrace slabs (measurement aid)
constexpr int kTrace = 1024;
__device__ long long g_trace[4][kTrace];     // [0] MMA issue (after tempty wait)  [1] epilogue saw tfull  [2] load landed  [3] TMA stage ready

template <int CG>
__global__ void __launch_bounds__(k2Threads, 1)
k_gemm4_tc2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const float *__restrict__ as, const float *__restrict__ bs, uint32_t M, uint32_t N, uint32_t K,
            float *__restrict__ c, uint64_t ldc, int dbg) {
    constexpr int k2Stages = PipeCfg<CG>::kStages, k2AStage = PipeCfg<CG>::kAStage, k2StageBytes = PipeCfg<CG>::kStageBytes;
    constexpr int kTileM = CG * k2BM;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem + k2Stages * k2StageBytes;
    const uint32_t full = bars, empty = full + 8 * k2Stages, tfull = empty + 8 * k2Stages, tempty = tfull + 32;
    const uint32_t slot = tempty + 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t pair = blockIdx.x / CG, npairs = gridDim.x / CG;
    const uint32_t tiles_m = (M + kTileM - 1) / kTileM, tiles_n = N / k2BN, ntiles = tiles_m * tiles_n;
    const uint32_t kblocks = K / k2BK, KB = K >> 6;

    if (threadIdx.x == 0) {
        for (int i = 0; i < k2Stages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * i), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * i), "r"(1) : "memory");
        }
        for (int b = 0; b < 4; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tfull + 8 * b), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tempty + 8 * b), "r"(CG * k2EpiWarps) : "memory");
        }
        mbar_fence_init();
        tma_prefetch_descriptor(&map_a);
        tma_prefetch_descriptor(&map_b);
    }
    if (warp == 2) {
        if (CG == 2) asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        else         asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));

    if (warp < 4) {
        reg_dealloc<32>();
        if (warp == 0) {
            // ===== TMA producer (both CTAs): own 128 rows of A', own 64 rows of B' =====
            const uint32_t full_leader = full & 0xFEFFFFFFu;           // same offset in the even CTA of the pair
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = pair; t < ntiles; t += npairs) {
                uint32_t tm, tn;
                tile_coords2(t, tiles_m, tiles_n, tm, tn);
                const int row_a = (int)(tm * kTileM + rank * k2BM), row_b = (int)(tn * k2BN + rank * (k2BN / 2));
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    mbar_wait_a(empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx_a(full + 8 * stage, CG * k2StageBytes);
                        const uint32_t sa = smem + stage * k2StageBytes;
                        if (CG == 2) {
                            tma_load_2d_2sm(sa, &map_a, (int)(kb * k2BK), row_a, full_leader + 8 * stage);
                            tma_load_2d_2sm(sa + k2AStage, &map_b, (int)(kb * k2BK), row_b, full_leader + 8 * stage);
                        } else {
                            tma_load_2d_a(sa, &map_a, (int)(kb * k2BK), row_a, full + 8 * stage);
                            tma_load_2d_a(sa + k2AStage, &map_b, (int)(kb * k2BK), row_b, full + 8 * stage);
                        }
                    }
                    __syncwarp();
                    if (++stage == k2Stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1 && rank == 0) {
            // ===== MMA issuer (leader CTA): slab g -> TMEM slot g & 3 =====
            const uint32_t idesc = umma_idesc(UMMA_E4M3, kTileM, k2BN);
            uint32_t stage = 0, phase = 0, g = 0;
            for (uint32_t t = pair; t < ntiles; t += npairs) {
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    mbar_wait_a(full + 8 * stage, phase);
                    tc_fence_after();
                    if ((dbg & 4) && blockIdx.x == 0 && g < kTrace) g_trace[3][g] = clock64();
                    const uint32_t sa = smem + stage * k2StageBytes;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + k2AStage);
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h, ++g) {
                        const uint32_t b = g & 3;
                        mbar_wait_a(tempty + 8 * b, ((g >> 2) & 1) ^ 1);
                        tc_fence_after();
                        if ((dbg & 4) && blockIdx.x == 0 && g < kTrace) g_trace[0][g] = clock64();
                        const uint32_t d = tmem + b * 128;
                        if (elect_one()) {
                            umma_ss<UMMA_E4M3, CG>(d, da + 4 * h, db + 4 * h, idesc, 0);
                            umma_ss<UMMA_E4M3, CG>(d, da + 4 * h + 2, db + 4 * h + 2, idesc, 1);
                            if (CG == 2) {
                                umma_commit_mc2(tfull + 8 * b);
                                if (h == 1) umma_commit_mc2(empty + 8 * stage);  // smem stage free once its MMAs retire
                            } else {
                                umma_commit_a<1>(tfull + 8 * b);
                                if (h == 1) umma_commit_a<1>(empty + 8 * stage);
                            }
                        }
                        __syncwarp();
                    }
                    if (++stage == k2Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: warps 4..19; warp = lane quadrant q x column quarter cq (32 columns) =====
        // Four warps per scheduler: while one waits for its tcgen05.ld to land, the others keep the fp32 pipe busy.
        // Each thread: 32 accumulators + two 32-register landing buffers (slab j is multiplied out of one while
        // slab j+1 lands in the other); a TMEM slot is released as soon as its load has landed, before its FMAs.
        reg_alloc<112>();
        const uint32_t q = warp & 3, cq = (uint32_t)(warp - 4) >> 2;
        const uint32_t taddr = tmem + ((q * 32) << 16) + cq * 32;
        uint32_t tempty_leader = tempty;
        if (CG == 2) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(tempty_leader) : "r"(tempty), "r"(0));
        uint64_t acc[16];
        uint32_t ra[32], rb[32];
        uint32_t g2 = 0;                 // slab-pair counter: pair p uses slots {0,1} (p even) or {2,3} (p odd), parity (p >> 1) & 1
        for (uint32_t t = pair; t < ntiles; t += npairs) {
            uint32_t tm, tn;
            tile_coords2(t, tiles_m, tiles_n, tm, tn);
            const uint32_t row0 = tm * kTileM + rank * k2BM;            // first row of this CTA's part of the tile
            const bool live = row0 < M;                                  // M is a multiple of 128: all in or all out
            const float *pa = as + (uint64_t)(live ? (row0 >> 6) + (q >> 1) : 0) * KB;
            const float *pb = bs + (uint64_t)(tn * 2 + (cq >> 1)) * KB;
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0ull;
            float sv = 0.f;
            {   // first slab of the tile -> ra
                const uint32_t hb = g2 & 1, par = (g2 >> 1) & 1;
                mbar_wait_a(tfull + 16 * hb, par);
                tc_fence_after();
                tmem_ld32(taddr + 256 * hb, ra);
                tmem_ld_wait(ra);
                tc_fence_before();
                if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty_leader + 16 * hb); else mbar_arrive_a(tempty_leader + 16 * hb); }
            }
            for (uint32_t kb = 0; kb < KB; kb += 2) {
                if ((kb & 31) == 0) {        // lane l owns the scale of slab kb + l: s = (sA * (1/49)) * sB
                    const uint32_t kl = min(kb + lane, KB - 1);
                    sv = __fmul_rn(__fmul_rn(__ldg(pa + kl), 1.0f / 49.0f), __ldg(pb + kl));
                }
                const float s_even = __shfl_sync(0xFFFFFFFFu, sv, kb & 31), s_odd = __shfl_sync(0xFFFFFFFFu, sv, (kb + 1) & 31);
                const uint32_t hb = g2 & 1, par = (g2 >> 1) & 1;
                // even slab is in ra; start the odd slab (slot 2hb+1) into rb, multiply, then complete + release it
                mbar_wait_a(tfull + 16 * hb + 8, par);
                tc_fence_after();
                if ((dbg & 4) && blockIdx.x == 0 && warp == 4 && 2 * g2 + 1 < kTrace) g_trace[1][2 * g2 + 1] = clock64();
                tmem_ld32(taddr + 256 * hb + 128, rb);
#pragma unroll
                for (int j = 0; j < 16; ++j) ffma2(acc[j], s_even, ra[2 * j], ra[2 * j + 1]);
                tmem_ld_wait(rb);
                tc_fence_before();
                if ((dbg & 4) && blockIdx.x == 0 && warp == 4 && 2 * g2 + 1 < kTrace) g_trace[2][2 * g2 + 1] = clock64();
                if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty_leader + 16 * hb + 8); else mbar_arrive_a(tempty_leader + 16 * hb + 8); }
                ++g2;
                const bool more = kb + 2 < KB;
                const uint32_t hn = g2 & 1, parn = (g2 >> 1) & 1;
                if (more) {                  // next pair's even slab (slot 2hn) into ra
                    mbar_wait_a(tfull + 16 * hn, parn);
                    tc_fence_after();
                    if ((dbg & 4) && blockIdx.x == 0 && warp == 4 && 2 * g2 < kTrace) g_trace[1][2 * g2] = clock64();
                    tmem_ld32(taddr + 256 * hn, ra);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) ffma2(acc[j], s_odd, rb[2 * j], rb[2 * j + 1]);
                if (more) {
                    tmem_ld_wait(ra);
                    tc_fence_before();
                    if ((dbg & 4) && blockIdx.x == 0 && warp == 4 && 2 * g2 < kTrace) g_trace[2][2 * g2] = clock64();
                    if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty_leader + 16 * hn); else mbar_arrive_a(tempty_leader + 16 * hn); }
                }
            }
            if (live) {
                float *crow = c + (uint64_t)(row0 + q * 32 + lane) * ldc + (uint64_t)tn * k2BN + cq * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o;
                    o.x = __uint_as_float((uint32_t)acc[2 * j]);     o.y = __uint_as_float((uint32_t)(acc[2 * j] >> 32));
                    o.z = __uint_as_float((uint32_t)acc[2 * j + 1]); o.w = __uint_as_float((uint32_t)(acc[2 * j + 1] >> 32));
                    *reinterpret_cast<float4 *>(crow + 4 * j) = o;
                }
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();    // nobody leaves while the peer's MMAs / remote arrivals may still touch this CTA
    else __syncthreads();
    if (warp == 2) tmem_dealloc<CG>(tmem, 512);
}

template <int CG>
static int launch_pipe(const uint8_t *a8, const float *as, const uint8_t *b8, const float *bs, uint64_t M, uint64_t N,
                       uint64_t K, float *c, uint64_t ldc, cudaStream_t stream) {
    CUtensorMap map_a, map_b;
    int rc = make_tensor_map_u8_2d_sw128(&map_a, a8, M, K, k2BM);
    if (rc != CLOVER_OK) return rc;
    rc = make_tensor_map_u8_2d_sw128(&map_b, b8, N, K, k2BN / CG);
    if (rc != CLOVER_OK) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    CLOVER_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        CLOVER_CUDA_CHECK(cudaFuncSetAttribute(k_gemm4_tc2<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, PipeCfg<CG>::kSmem));
        attr_set[dev & 63] = true;
    }
    const uint64_t ntiles = ((M + CG * k2BM - 1) / (CG * k2BM)) * (N / k2BN);
    const unsigned groups = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sm_count() / CG);
    static const int dbg = [] { const char *e = getenv("CLOVER_GEMM_DBG"); return e ? atoi(e) : 0; }();   // 4: hand-off trace of CTA 0 to stderr
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CG * groups);
    cfg.blockDim = dim3(k2Threads);
    cfg.dynamicSmemBytes = PipeCfg<CG>::kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CLOVER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gemm4_tc2<CG>, map_a, map_b, as, bs, (uint32_t)M, (uint32_t)N, (uint32_t)K, c, ldc, dbg));
    count_launch();
    if (dbg & 4) {
        static long long h[4][kTrace];
        CLOVER_CUDA_CHECK(cudaDeviceSynchronize());
        CLOVER_CUDA_CHECK(cudaMemcpyFromSymbol(h, g_trace, sizeof(h)));
        double per = 0, exec = 0, land = 0, rel = 0, tma = 0; int n = 0;
        for (int g = 260; g + 4 < 500; ++g, ++n) {          // second tile of CTA 0: steady state
            per += (double)(h[0][g + 1] - h[0][g]);
            exec += (double)(h[1][g] - h[0][g]);             // MMA issue -> epilogue sees tfull
            land += (double)(h[2][g] - h[1][g]);             // tcgen05.ld round trip
            rel += (double)(h[0][g + 4] - h[2][g]);          // landed(g) -> MMA issue of slab g+4 (same slot)
            tma += (double)(h[0][g & ~1] - h[3][g & ~1]);    // stage ready -> first MMA of the stage issued
        }
        fprintf(stderr, "[gemm trace CG=%d] per slab: issue period %.0f | issue->tfull seen %.0f | ld round trip %.0f | landed->issue(g+4) %.0f | "
                "stage ready->issue %.0f cycles\n", CG, per / n, exec / n, land / n, rel / n, tma / n);
        for (int g = 300; g < 312; ++g)
            fprintf(stderr, "   g=%d issue %lld  seen +%lld  landed +%lld  stage-ready %+lld\n", g, h[0][g] - h[0][300], h[1][g] - h[0][g], h[2][g] - h[0][g],
                    h[3][g & ~1] - h[0][g]);
    }
    return launch_status("k_gemm4_tc2");
}

int gemm4_tc2_expanded(const uint8_t *a8, const float *as, const uint8_t *b8, const float *bs, uint64_t M, uint64_t N,
                       uint64_t K, float *c, uint64_t ldc, cudaStream_t stream, int cta_group) {
    return cta_group == 2 ? launch_pipe<2>(a8, as, b8, bs, M, N, K, c, ldc, stream) : launch_pipe<1>(a8, as, b8, bs, M, N, K, c, ldc, stream);
}

}  // namespace clover
