// gemm4_tc.cu - 4-bit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[i][j] = sum_kb (sA[i>>6][kb] * (1/49) * sB[j>>6][kb]) * I_kb[i][j]        (gemm4.cu)
//
// The scale changes every K-slab of 64, so the integer partial I_kb of every slab has to leave the tensor
// core, be multiplied by its fp32 scale and be accumulated in fp32 - one FMA per output element per slab.
// That FMA stream (M*N*K/64 of them, as many cycles on the fp32 pipe as the slab's MMAs take on the tensor
// pipe) is the real bound of this kernel, so everything else is kept off the CUDA cores:
//
//   1. k_expand_e4m3 (one HBM pass, declared: +0.75 GiB of traffic at 16384^3, a few % of the GEMM):
//      two's-complement nibbles -> FP8 E4M3 bytes. Integers -8..8 are exact in E4M3 and
//      tcgen05.mma kind::f8f6f4 accumulates in fp32, where slab sums (<= 64*49) are exact integers:
//      the accumulator that comes out of TMEM already IS float(I_kb) - no int->float conversion in the
//      epilogue (kind::i8 would cost two extra fp32-pipe operations per element). Exactness of this
//      path was verified on B200 against the integer product (tools/mma_probe.cu check).
//   2. k_gemm4_tc: persistent, one CTA per SM, 128x256 output tile, warp-specialised:
//        warp 0      TMA producer: A' 128x128 B and B' 256x128 B boxes (SWIZZLE_128B) into a 4-stage ring
//        warp 1      MMA issuer: per stage 2 slabs x 2 tcgen05.mma (M128 N256 K32) into one of two
//                    256-column TMEM buffers, tcgen05.commit -> mbarrier per slab
//        warp 2      TMEM allocator
//        warps 4-19  epilogue: warp = 32 rows x 64 columns (one scale tile), thread = one row, 64 fp32
//                    accumulators in registers for the whole K loop; per slab 2 x tcgen05.ld.32x32b.x32 and
//                    32 packed fma.rn.f32x2 with the slab's scale (s = (sA*(1/49))*sB, the reference's order);
//                    four warps per scheduler hide the TMEM-load and mbarrier latencies of one another.
//      fp32 accumulation is sequential in kb - identical to k_gemm4_simt, so the two kernels agree bit for bit.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "runtime.cuh"
#include "tcgen05.cuh"

namespace clover {

// ---------------------------------------------------------------------------------------------------------
// nibble -> E4M3 expansion
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// w holds 8 nibbles of the reference layout (byte i: element 2i in the HIGH nibble). o0 = elements 0..3,
// o1 = elements 4..7 as E4M3 bytes. All 16 codes are handled (-8 -> 0xD0), although quantize never emits -8.
__device__ __forceinline__ void nib8_to_e4m3(uint32_t w, uint32_t &o0, uint32_t &o1) {
    const uint32_t x = ((w & 0x0F0F0F0Fu) << 4) | ((w >> 4) & 0x0F0F0F0Fu);   // nibble j = element j
    const uint32_t idx = x & 0x77777777u;
    // byte k of the tables: E4M3(k) and E4M3(k - 8)
    const uint32_t PL = 0x44403800u, PH = 0x4E4C4A48u, NL = 0xCACCCED0u, NH = 0xB8C0C4C8u;
    const uint32_t p0 = prmt(PL, PH, idx), p1 = prmt(PL, PH, idx >> 16);
    const uint32_t n0 = prmt(NL, NH, idx), n1 = prmt(NL, NH, idx >> 16);
    const uint32_t xs = x << 4;                                                  // msb of byte k = sign of nibble 2k
    const uint32_t m0 = prmt(xs, x, 0xD9C8u), m1 = prmt(xs, x, 0xFBEAu);        // 0xFF where the nibble is negative
    o0 = (p0 & ~m0) | (n0 & m0);
    o1 = (p1 & ~m1) | (n1 & m1);
}

__global__ void __launch_bounds__(256) k_expand_e4m3(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = ldg_stream(in + i);
        uint4 a, b;
        nib8_to_e4m3(w.x, a.x, a.y);
        nib8_to_e4m3(w.y, a.z, a.w);
        nib8_to_e4m3(w.z, b.x, b.y);
        nib8_to_e4m3(w.w, b.z, b.w);
        out[2 * i] = a;
        out[2 * i + 1] = b;
    }
}

// ---------------------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBN = 256, kBK = 128, kStages = 4;
constexpr int kAStage = kBM * kBK, kBStage = kBN * kBK, kStageBytes = kAStage + kBStage;
constexpr int kEpiWarps = 16;                      // 4 lane quadrants x 4 column blocks of 64
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr uint32_t kGroupM = 16;     // tiles are walked in groups of 16 tile-rows so that concurrent CTAs share A and B' panels in L2

template <uint32_t GROUP = kGroupM>
__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t tiles_m, uint32_t tiles_n, uint32_t &tm, uint32_t &tn) {
    const uint32_t group_sz = GROUP * tiles_n;
    const uint32_t g = t / group_sz, r = t % group_sz;
    const uint32_t first = g * GROUP;
    const uint32_t gm = min(GROUP, tiles_m - first);
    tm = first + r % gm;
    tn = r / gm;
}

// one K-slab of one 32-lane x 64-column block: acc += s * D, D read from TMEM in two 32-column chunks
__device__ __forceinline__ void epilogue_slab(uint64_t *acc, uint32_t *r, float s, uint32_t taddr, uint32_t tfull, uint32_t tempty,
                                              uint32_t parity, int lane) {
    mbar_wait_a(tfull, parity);
    tc_fence_after();
    tmem_ld32(taddr, r);
    tmem_ld_wait(r);
#pragma unroll
    for (int j = 0; j < 16; ++j) ffma2(acc[j], s, r[2 * j], r[2 * j + 1]);
    tmem_ld32(taddr + 32, r);
    tmem_ld_wait(r);
    tc_fence_before();                                   // this warp's share of the buffer is drained
    if (lane == 0) mbar_arrive_a(tempty);
#pragma unroll
    for (int j = 0; j < 16; ++j) ffma2(acc[16 + j], s, r[2 * j], r[2 * j + 1]);
}

// TMA load with an L2 eviction-priority hint
__device__ __forceinline__ void tma_load_2d_hint(uint32_t smem_dst, const void *tensor_map, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_dst), "l"(tensor_map), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm4_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
           const float *__restrict__ as, const float *__restrict__ bs, uint32_t M, uint32_t N, uint32_t K,
           float *__restrict__ c, uint64_t ldc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;             // 1024-aligned: SWIZZLE_128B atoms
    // barrier block behind the stages: full[4] empty[4] tfull[2] tempty[2] tmem_slot
    const uint32_t bars = smem + kStages * kStageBytes;
    const uint32_t full = bars, empty = bars + 8 * kStages, tfull = bars + 16 * kStages, tempty = tfull + 32;
    const uint32_t slot = tempty + 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_n = (N + kBN - 1) / kBN;
    const uint32_t tiles_m = M / kBM, ntiles = tiles_m * tiles_n;
    const uint32_t kblocks = K / kBK, KB = K >> 6, NB = N >> 6;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * i), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * i), "r"(1) : "memory");
        }
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tfull + 8 * b), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tempty + 8 * b), "r"(kEpiWarps) : "memory");
        }
        mbar_fence_init();
        tma_prefetch_descriptor(&map_a);
        tma_prefetch_descriptor(&map_b);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));

    if (warp < 4) {
        reg_dealloc<32>();   // 128*32 + 512*112 = 640*96: exactly the registers this CTA was launched with
        if (warp == 0) {
            // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
            // The persistent grid walks the tiles in groups of kGroupM tile-rows: within a group the 16 A' panels (2 MiB each
            // at K = 16384) are re-read by every wave while the B' panels stream through once. A' is therefore loaded with an
            // evict-last hint and B' with evict-first, so that the streaming operand does not push the re-used one out of L2
            // (round 1: 4.16 GB of DRAM reads per 16384^3 launch against 0.54 GB of operands).
            uint64_t keep, stream;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                uint32_t tm, tn;
                tile_coords<kGroupM>(t, tiles_m, tiles_n, tm, tn);
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    mbar_wait_a(empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx_a(full + 8 * stage, kStageBytes);
                        const uint32_t sa = smem + stage * kStageBytes;
                        tma_load_2d_hint(sa, &map_a, (int)(kb * kBK), (int)(tm * kBM), full + 8 * stage, keep);
                        tma_load_2d_hint(sa + kAStage, &map_b, (int)(kb * kBK), (int)(tn * kBN), full + 8 * stage, stream);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer: per stage, K-slab 0 -> TMEM buffer 0, K-slab 1 -> buffer 1 =====
            const uint32_t idesc = umma_idesc(UMMA_E4M3, kBM, kBN);
            uint32_t stage = 0, phase = 0, pair = 0;
            for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                for (uint32_t kb = 0; kb < kblocks; ++kb, ++pair) {
                    mbar_wait_a(full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem + stage * kStageBytes;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + kAStage);
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {
                        mbar_wait_a(tempty + 8 * h, (pair & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem + h * 256;
                        if (elect_one()) {
                            umma_ss<UMMA_E4M3, 1>(d, da + 4 * h, db + 4 * h, idesc, 0);
                            umma_ss<UMMA_E4M3, 1>(d, da + 4 * h + 2, db + 4 * h + 2, idesc, 1);
                            umma_commit_a<1>(tfull + 8 * h);
                            if (h == 1) umma_commit_a<1>(empty + 8 * stage);      // smem stage free once its MMAs retire
                        }
                        __syncwarp();
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: warps 4..19. Warp = lane quadrant q (TMEM lanes 32q..32q+31 = tile rows) x column block cb =====
        reg_alloc<112>();
        const uint32_t q = warp & 3, cb = (uint32_t)(warp - 4) >> 2;
        const uint32_t taddr = tmem + ((q * 32) << 16) + cb * 64;
        uint64_t acc[32];
        uint32_t r[32];
        uint32_t pair = 0;
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            uint32_t tm, tn;
            tile_coords<kGroupM>(t, tiles_m, tiles_n, tm, tn);
            const uint32_t jb = tn * 4 + cb;
            const bool live = jb < NB;                      // N is a multiple of 128: a 64-column block is all in or all out
            const float *pa = as + (uint64_t)(tm * 2 + (q >> 1)) * KB;
            const float *pb = bs + (uint64_t)(live ? jb : 0) * KB;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0ull;
            for (uint32_t kb0 = 0; kb0 < KB; kb0 += 32) {
                // lane l owns the scale of slab kb0 + l: s = (sA * (1/49)) * sB, the reference's order
                const uint32_t kl = min(kb0 + lane, KB - 1);
                const float sv = __fmul_rn(__fmul_rn(__ldg(pa + kl), 1.0f / 49.0f), __ldg(pb + kl));
                const uint32_t n = min(32u, KB - kb0);
                for (uint32_t k = 0; k < n; k += 2, ++pair) {
                    const float s_even = __shfl_sync(0xFFFFFFFFu, sv, k), s_odd = __shfl_sync(0xFFFFFFFFu, sv, k + 1);
                    epilogue_slab(acc, r, s_even, taddr, tfull, tempty, pair & 1, lane);
                    epilogue_slab(acc, r, s_odd, taddr + 256, tfull + 8, tempty + 8, pair & 1, lane);
                }
            }
            if (live) {
                float *crow = c + (uint64_t)(tm * kBM + q * 32 + lane) * ldc + (uint64_t)tn * kBN + cb * 64;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float4 o;
                    o.x = __uint_as_float((uint32_t)acc[2 * j]);     o.y = __uint_as_float((uint32_t)(acc[2 * j] >> 32));
                    o.z = __uint_as_float((uint32_t)acc[2 * j + 1]); o.w = __uint_as_float((uint32_t)(acc[2 * j + 1] >> 32));
                    __stcs(reinterpret_cast<float4 *>(crow + 4 * j), o);          // C is written once and not re-read: streaming store
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// workspace for the expanded operands, keyed by (device, stream)
static int gemm_workspace(cudaStream_t stream, size_t bytes, void **out) {
    return stream_scratch(kScratchGemm, stream, bytes, 0, out);
}

int gemm4_expand(const int8_t *values, uint64_t rows, uint64_t cols, uint8_t *out, cudaStream_t stream) {
    const uint64_t n16 = rows * cols / 32;       // 16-byte input chunks
    if (n16 == 0) return CLOVER_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>((n16 + 255) / 256, (uint64_t)sm_count() * 16);
    k_expand_e4m3<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(values), reinterpret_cast<uint4 *>(out), n16);
    count_launch();
    return launch_status("k_expand_e4m3");
}

int gemm4_tc_expanded(const uint8_t *a8, const float *as, const uint8_t *b8, const float *bs, uint64_t M, uint64_t N,
                      uint64_t K, float *c, uint64_t ldc, cudaStream_t stream) {
    CUtensorMap map_a, map_b;
    int rc = make_tensor_map_u8_2d_sw128(&map_a, a8, M, K, kBM);
    if (rc != CLOVER_OK) return rc;
    rc = make_tensor_map_u8_2d_sw128(&map_b, b8, N, K, kBN);
    if (rc != CLOVER_OK) return rc;
    CLOVER_CUDA_CHECK(cudaFuncSetAttribute(k_gemm4_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    const uint64_t ntiles = (M / kBM) * ((N + kBN - 1) / kBN);
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sm_count());
    k_gemm4_tc<<<grid, kGemmThreads, kGemmSmem, stream>>>(map_a, map_b, as, bs, (uint32_t)M, (uint32_t)N, (uint32_t)K, c, ldc);
    count_launch();
    return launch_status("k_gemm4_tc");
}

int gemm4_tc(const int8_t *av, const float *as, const int8_t *btv, const float *bts, uint64_t M, uint64_t N, uint64_t K,
             float *c, uint64_t ldc, cudaStream_t stream) {
    void *ws = nullptr;
    int rc = gemm_workspace(stream, (M + N) * K, &ws);
    if (rc != CLOVER_OK) return rc;
    uint8_t *a8 = static_cast<uint8_t *>(ws), *b8 = a8 + M * K;
    rc = gemm4_expand(av, M, K, a8, stream);
    if (rc != CLOVER_OK) return rc;
    rc = gemm4_expand(btv, N, K, b8, stream);
    if (rc != CLOVER_OK) return rc;
    return gemm4_tc_expanded(a8, as, b8, bts, M, N, K, c, ldc, stream);
}

}  // namespace clover
