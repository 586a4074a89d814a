// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] . B[N,K]^T, bf16 operands (K-major), fp32 accumulation in TMEM.
//
// One CTA computes one 128 x BN output tile:
//   warp 0 (one lane)  : TMA producer  -- cp.async.bulk.tensor.2d of the A (128x64) and B (BNx64) K-blocks, 128B swizzle,
//                        rows beyond M are zero-filled by the tensor map (no guards in the kernel)
//   warp 1 (one lane)  : MMA issuer    -- tcgen05.mma.cta_group::1.kind::f16, 4 x (K=16) per K-block, accumulator in TMEM;
//                        tcgen05.commit releases each smem stage and finally signals the epilogue
//   warps 2..5         : TMEM -> registers (tcgen05.ld 32x32b) -> fp32 tile in shared memory
//   all 6 warps        : the same epilogue functors as the other mainloops (EpiStore / EpiGradStore / EpiLstmFwd / EpiLstmBwd)
// Every mbarrier wait is bounded (trap after ~2 s) so a protocol bug cannot hang the GPU.
#pragma once
#include <cuda.h>

#include <unordered_map>

#include "common.cuh"

namespace tc {

// Debug probe (s2vt_debug_probe): when set, CTA (0,0) of every tcgen05 GEMM launch records %globaltimer at its phase
// boundaries into slot [launch*8 .. launch*8+7] of the buffer (slot 0 of the buffer is the launch counter).
static __device__ unsigned long long* g_probe = nullptr;   // static: one copy per translation unit (s2vt_debug_probe arms the S2VT engine's)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

constexpr int BM = 128, BK = 64, NTHREADS = 192;
constexpr uint32_t FMT_A_F16 = 1u << 7, FMT_B_F16 = 1u << 10;
template <typename T> struct FmtOf { static constexpr uint32_t A = 0, B = 0; };
template <> struct FmtOf<f16> { static constexpr uint32_t A = FMT_A_F16, B = FMT_B_F16; };   // NTHREADS: staged-epilogue kernels (2 role warps + 4 epilogue warps)

template <int BN_, int NTHREADS_ = tc::NTHREADS>
struct Cfg {
    static constexpr int BM = tc::BM, BN = BN_, BK = tc::BK, NTHREADS = NTHREADS_;
    static constexpr int LDC = BN + 4;
    // BN <= 64 (per-step GEMMs): small enough that CTAs of the NEXT step's kernel (programmatic dependent launch) fit on the
    // SM beside the running ones: BN=32 -> 100 KB (2 per SM), BN=64 -> 72 KB (3 per SM)
    static constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 5 : (BN >= 64 ? 3 : 5));
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int EPI_BYTES = BM * LDC * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) + BAR_BYTES + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;   // power of two >= 32
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6), a/b_format BF16 [7,10)/[10,13),
    // a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).  Format 0 = F16, 1 = BF16: the kernels take a run-time
    // `fmt` word whose bits are CLEARED from IDESC -- FMT_A_F16 / FMT_B_F16 turn an operand into IEEE fp16 (same tensor maps: TMA
    // moves 16-bit payloads), so the fp16 x fp16 forward products and the bf16 x bf16 backward products share one kernel.  Both
    // operands must have the SAME format: a mixed descriptor raises an illegal-instruction fault on B200 (measured).
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the warp in hardware until the phase completes or a time limit passes, so the loop below wakes rarely; the
// watchdog clock is read only every 256 wake-ups (a clock64 + compare per poll in the tcgen05.mma issuing thread and in the
// epilogue warps spinning beside it cost more than the MMAs themselves for narrow tiles).
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    return done != 0;
}
static __device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
    long long t0 = clock64();
    for (unsigned spins = 1;; ++spins) {
        if (mbar_try(addr, parity)) return;
        if ((spins & 255u) == 0 && clock64() - t0 > 4000000000LL) {
            printf("s2vt: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try(addr, parity)) return;
    mbar_wait_slow(addr, parity);
}
// One lane of a converged warp.  The role loops below run on the WHOLE warp and only the tcgen05.mma / TMA / commit instructions are
// predicated on the elected lane: descriptors and coordinates are then computed in warp-uniform code and stay in uniform registers.
// Issued from inside `if (lane == 0)` the compiler had to move every operand into uniform registers with an ELECT / R2UR loop
// before each UTCHMMA -- about as expensive as the 64-cycle MMA itself for tiles of N <= 128.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d_raw(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void tma_load_3d_raw(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#ifdef S2VT_TMA_SPLIT
// experiment: the tensor maps are built with 64-row boxes and every tile is fetched as rows/64 separate TMA instructions
__device__ int g_split_rows_a, g_split_rows_b;
#endif
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer, int rows = 0) {
#ifdef S2VT_TMA_SPLIT
    for (int r = 0; r < rows; r += 64) tma_load_2d_raw((char*)smem_dst + r * 128, map, bar, c_inner, c_outer + r);
#else
    tma_load_2d_raw(smem_dst, map, bar, c_inner, c_outer);
#endif
}
// K-major, 128-byte swizzle smem matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO=1 [16,30),
// SBO = 1024 B >> 4 = 64 [32,46) (stride between 8-row groups), version 1 at [46,48), SWIZZLE_128B = 2 at [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, 128-byte swizzle: each K index is a 128-byte row of 64 MN-contiguous elements (what a TMA box {64 features,
// 64 rows} of a row-major [rows, features] matrix produces); 8 K rows form a 1024-byte atom (SBO), the next 64 MN
// elements live one box (8192 bytes) further (LBO).  cute::UMMA canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (512ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "h"(mask) : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One 32-column accumulator chunk through a direct epilogue.  EpiLstmFwd: 32 packed gate columns = one 8-unit call.
// EpiLstmBwd (kUnitsPerChunk = 8): the 32 columns are 32 units, handled as four 8-unit calls; the operands of the next
// call are fetched before the current one is computed.
template <class Epi, class = void> struct DirectTraits { static constexpr int kSub = 32; };
template <class Epi> struct DirectTraits<Epi, decltype((void)Epi::kUnitsPerChunk)> { static constexpr int kSub = Epi::kUnitsPerChunk; };
template <class Epi>
__device__ __forceinline__ void direct_chunk(const typename Epi::Params& ep, int gr, int gc, const float* v, typename Epi::Pre& pre, bool more) {
    constexpr int SUB = DirectTraits<Epi>::kSub;
    if constexpr (SUB == 32) {
        Epi::direct(ep, gr, gc, v, pre);
        if (more) Epi::prefetch(ep, gr, gc + 32, pre);
    } else {
#pragma unroll
        for (int s = 0; s < 32; s += SUB) {
            typename Epi::Pre cur = pre;
            if (s + SUB < 32 || more) Epi::prefetch(ep, gr, gc + s + SUB, pre);
            Epi::direct(ep, gr, gc + s, v + s, cur);
        }
    }
}

// Epilogue warps: the staged epilogues use 4 (one per TMEM lane quarter); the direct (register) epilogues use one warp per
// (lane quarter, 32-column chunk) so every thread handles exactly one chunk whose operands were prefetched.
template <class Epi, class = void> struct StagedWarps { static constexpr int value = 4; };
template <class Epi> struct StagedWarps<Epi, decltype((void)Epi::kEpiWarps)> { static constexpr int value = Epi::kEpiWarps; };   // compute-heavy staged epilogues
template <int BN, class Epi> struct Threads {
    static constexpr int NEPI = Epi::kDirect ? 4 * (BN / 32) : StagedWarps<Epi>::value;
    static constexpr int N = 64 + 32 * NEPI;
};

__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// KS > 1: split-K over a cluster of KS CTAs (cluster dims (1,1,KS), rank = blockIdx.z).  Each CTA accumulates its K range
// in TMEM; partial tiles are exchanged through distributed shared memory so that CTA r reduces and finishes rows
// [32r, 32r+32) of the tile (KS == 4): the K loop AND the epilogue are both split four ways.
// CX x CY > 1: TMA multicast over a (CX, CY, 1) cluster.  The CX CTAs of a cluster row compute the same 128 output rows, so
// each loads 128/CX rows of the A K-block and multicasts them to the row; the CY CTAs of a cluster column share the B tile
// the same way.  L2 reads per CTA drop from (128 + BN) to (128/CX + BN/CY) rows per K-block.  A stage may only be refilled
// when every CTA that receives part of it has consumed it: the MMA warp's tcgen05.commit arrives (multicast) on the
// `empty` barrier of each CTA that sends to it, and `empty` counts CX + CY - 1 arrivals.
// MN == true: both operands are MN-major -- C[Mf, Nf] = sum_r X[r, Mf] . Y[r, Nf] for row-major X, Y (the weight-gradient
// products), so no transposed copies are needed; the contraction runs over rows and its tail is zero-filled by TMA.
// Accumulator tiles per output tile.  Splitting the K loop over several independent TMEM accumulators was tried against the
// hypothesis that back-to-back tcgen05.mma into one tile wait for each other: scripts/micro/mma_rate.cu shows they do not (a
// 128 x N x 16 MMA costs ~64 cycles for every N <= 128, dependent or not), while every extra tile costs a full TMEM read in the
// epilogue (64 B/clk: 0.5 us per 128 x 128 tile).  Kept at 1; the plumbing stays for experiments.
// Round 2: the 32-column tiles (64-row chains and per-step kernels) use TWO accumulators, even K-blocks into the first and odd ones into the second, summed by
// the epilogue -- so that the 64-row chains can issue from two warps (gemm_tcgen05_chain.cuh, MMA2) and still sum in the order of the per-step kernels
// (bit-identical results across the variants).  Measured: 64-row forward chain 5.46 -> 4.90 us per step, 64-row BPTT chain 7.76 -> 7.2; for the 128-column
// tiles of the 320-row chains (fill- / exchange-bound) and the M = 64 halves of the pipelined chain a second issuing warp changed nothing or cost time.
// (four warps / four accumulators measured slower than two: 8.75 vs 8.67 ms per iteration)
template <int BN> struct ChainAcc { static constexpr int N = BN == 32 ? 2 : 1; };

template <int BN, class Epi, int KS, int CX = 1, int CY = 1, bool MN = false>
__global__ void __launch_bounds__(Threads<BN, Epi>::N) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                      int K, int a_rows, uint32_t fmt, typename Epi::Params ep) {
    // a_rows: rows of the A K-block actually fetched (TMA box height).  Problems with M <= 64 fetch 64 rows only; the rest of
    // the 128-row smem tile is stale and only feeds accumulator rows >= M, which no epilogue reads.
    using C = Cfg<BN, Threads<BN, Epi>::N>;
    static_assert(KS == 1 || (KS == 4 && Epi::kDirect), "split-K needs a direct epilogue and a cluster of 4");
    static_assert(KS == 1 || (CX == 1 && CY == 1), "split-K and multicast clusters are exclusive");
    static_assert(!MN || (KS == 1 && CX == 1 && CY == 1), "MN-major operands: plain kernel only");
    const uint32_t IDESC = (C::IDESC | (MN ? ((1u << 15) | (1u << 16)) : 0u)) & ~fmt;
    constexpr bool MC = CX * CY > 1;
    constexpr int NACC = ChainAcc<BN>::N;
    constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation: power of two <= 512 columns");
    constexpr int A_ROWS = BM / CX, B_ROWS = BN / CY;   // rows this CTA fetches of each tile
    const uint32_t stage_tx = (uint32_t)((MC || MN) ? C::STAGE_BYTES : a_rows * 128 + C::B_BYTES);
    uint32_t crank = 0;
    if constexpr (MC) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int cx = crank % CX, cy = crank / CX;
    const uint16_t row_mask = (uint16_t)(((1u << CX) - 1u) << (cy * CX));
    uint16_t col_mask = 0;
#pragma unroll
    for (int j = 0; j < CY; ++j) col_mask |= (uint16_t)(1u << (j * CX + cx));
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms need 1024 B alignment
    constexpr int MAIN = C::PIPE_BYTES > C::EPI_BYTES ? C::PIPE_BYTES : C::EPI_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + MAIN);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tmem_full = empty + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* Cs = reinterpret_cast<float*>(smem);   // aliases the pipeline buffers once every MMA has retired
    float* recv = reinterpret_cast<float*>(smem + MAIN + C::BAR_BYTES);   // [KS][32][BN] partial tiles from the cluster (KS > 1)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * C::BM, n0 = blockIdx.x * BN;
    const int rank = KS > 1 ? (int)blockIdx.z : 0;
    const int KBL = MN ? (K + C::BK - 1) / C::BK : K / C::BK / KS, kb0 = rank * KBL;   // this CTA's K-blocks (MN: ragged tail is zero-filled)
    __shared__ unsigned long long* probe;
    if (threadIdx.x == 0) {
        probe = nullptr;
        if (g_probe && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            unsigned long long slot = atomicAdd(g_probe, 1ull);
            if (slot < 4000) { probe = g_probe + 8 * (slot + 1); probe[0] = gtimer(); probe[6] = ((unsigned long long)(BN + 1000 * KS) << 32) | (unsigned)K; probe[7] = gridDim.x * gridDim.y * gridDim.z; }
        }
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CX + CY - 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-wide; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nacc = KBL < NACC ? KBL : NACC;           // accumulator tiles this launch writes
    if constexpr (KS > 1 || MC) cluster_sync_all();   // every CTA of the cluster is running and its barriers are initialised

    // Programmatic dependent launch: let the next kernel in the stream start its prologue now; the weight (B) tiles of the
    // first stages never depend on the previous kernel, so they are requested before waiting for it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int pre = KBL < C::STAGES ? KBL : C::STAGES;
    if (warp == 0 && elect_one()) {
        for (int i = 0; i < pre; ++i) {
            mbar_expect_tx(full + i, stage_tx);
            unsigned char* b = smem + i * C::STAGE_BYTES + C::A_BYTES + cy * B_ROWS * 128;
            if constexpr (MN) {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) tma_load_2d_raw(b + j * 8192, &mapB, full + i, n0 + 64 * j, (kb0 + i) * C::BK);
            }
            else if constexpr (CY > 1) tma_load_2d_mc(b, &mapB, full + i, (kb0 + i) * C::BK, n0 + cy * B_ROWS, col_mask);
            else tma_load_2d(b, &mapB, full + i, (kb0 + i) * C::BK, n0, BN);
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (probe && threadIdx.x == 0) probe[1] = gtimer();

    if (warp == 0) {
        const bool leader = elect_one();
        for (int i = 0; i < KBL; ++i) {
            const int s = i % C::STAGES;
            unsigned char* a = smem + s * C::STAGE_BYTES;
            if (i >= C::STAGES) {
                mbar_wait(empty + s, ((i / C::STAGES) - 1) & 1);
                if (leader) {
                    mbar_expect_tx(full + s, stage_tx);
                    unsigned char* b = a + C::A_BYTES + cy * B_ROWS * 128;
                    if constexpr (MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_2d_raw(b + j * 8192, &mapB, full + s, n0 + 64 * j, (kb0 + i) * C::BK);
                    }
                    else if constexpr (CY > 1) tma_load_2d_mc(b, &mapB, full + s, (kb0 + i) * C::BK, n0 + cy * B_ROWS, col_mask);
                    else tma_load_2d(b, &mapB, full + s, (kb0 + i) * C::BK, n0, BN);
                }
            }
            if (leader) {
                if constexpr (MN) {
                    tma_load_2d_raw(a, &mapA, full + s, m0, (kb0 + i) * C::BK);
                    tma_load_2d_raw(a + 8192, &mapA, full + s, m0 + 64, (kb0 + i) * C::BK);
                }
                else if constexpr (CX > 1) tma_load_2d_mc(a + cx * A_ROWS * 128, &mapA, full + s, (kb0 + i) * C::BK, m0 + cx * A_ROWS, row_mask);
                else tma_load_2d(a, &mapA, full + s, (kb0 + i) * C::BK, m0, a_rows);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const bool leader = elect_one();
        for (int i = 0; i < KBL; ++i) {
            const int s = i % C::STAGES;
            mbar_wait(full + s, (i / C::STAGES) & 1);
            if (i == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (probe && i == 0 && leader) probe[2] = gtimer();
            const uint32_t a = smem_u32(smem + s * C::STAGE_BYTES);
            const uint64_t adesc = MN ? make_desc_mn(a) : make_desc(a), bdesc = MN ? make_desc_mn(a + C::A_BYTES) : make_desc(a + C::A_BYTES);
            constexpr int KADV = MN ? 128 : 2;      // K-major: +32 bytes per K=16 step inside the swizzle atom; MN-major: +16 rows = 2048 bytes
            if (leader) {
#pragma unroll
                for (int k = 0; k < C::BK / 16; ++k)
                    mma_bf16(tmem_base + (uint32_t)((i % NACC) * BN), adesc + KADV * k, bdesc + KADV * k, IDESC, i >= NACC || k != 0);
                if constexpr (MC) mma_commit_mc(empty + s, (uint16_t)(row_mask | col_mask));
                else mma_commit(empty + s);             // implies tcgen05.fence::before_thread_sync
            }
        }
        if (leader) {
            mma_commit(tmem_full);
            if (probe) probe[3] = gtimer();
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int q = warp & 3;                         // a warp may only touch TMEM lanes [32 (warp % 4), +32)
        const int row = q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        if constexpr (Epi::kDirect && KS == 1) {
            // register epilogue: one 32-column chunk per thread; its operands are fetched while the MMAs are still running
            const int c0 = (e >> 2) * 32;
            typename Epi::Pre pre;
            Epi::prefetch(ep, m0 + row, n0 + c0, pre);
            mbar_wait(tmem_full, 0);
            if (probe && threadIdx.x == 64) probe[4] = gtimer();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[32];
            tmem_ld32(trow + (uint32_t)c0, v);
            for (int a = 1; a < nacc; ++a) {
                float w[32];
                tmem_ld32(trow + (uint32_t)(a * BN + c0), w);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += w[j];
            }
            direct_chunk<Epi>(ep, m0 + row, n0 + c0, v, pre, false);
        } else if constexpr (Epi::kDirect) {
            // split-K: final mapping of this thread = row (32 rank + t / (BN/8)), units [8 (t % (BN/8)), +8)
            constexpr int UPR = BN / 8;                 // 8-unit tasks per row
            const int t = threadIdx.x - 64;
            const int frow = t / UPR, fc8 = t % UPR;
            typename Epi::Pre pre;
            Epi::prefetch(ep, m0 + 32 * rank + frow, n0 + 8 * fc8, pre, 1);
            mbar_wait(tmem_full, 0);
            if (probe && threadIdx.x == 64) probe[4] = gtimer();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {   // send this warp's 32 x 32 partial (rows of lane quarter q) to CTA q, slot `rank`, 16-byte chunks XOR-swizzled by row
                const int c0 = (e >> 2) * 32;
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                for (int a = 1; a < nacc; ++a) {
                    float w[32];
                    tmem_ld32(trow + (uint32_t)(a * BN + c0), w);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += w[j];
                }
                // exchange layout [source rank][16-byte column group][row]: one store instruction of a warp covers 512 contiguous bytes
                const uint32_t base = cluster_map(smem_u32(recv), (uint32_t)q) + (uint32_t)(((rank * (BN / 4) + (c0 >> 2)) * 32) * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j)      // row slot XOR-ed with the 8-unit group index: the reducer's 16 lanes of a row then spread over the banks
                        st_cluster_f4(base + j * 512 + ((lane ^ ((((c0 >> 2) + j) >> 1) & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            Epi::prefetch(ep, m0 + 32 * rank + frow, n0 + 8 * fc8, pre, 2);
            asm volatile("barrier.cluster.arrive.release;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int src = 0; src < KS; ++src) {
                const float4* rp = reinterpret_cast<const float4*>(recv) + (src * (BN / 4) + 2 * fc8) * 32 + (frow ^ (fc8 & 7));
                const float4 x0 = rp[0], x1 = rp[32];
                acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w; acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
            }
            Epi::direct(ep, m0 + 32 * rank + frow, n0 + 8 * fc8, acc, pre);
        } else {
            mbar_wait(tmem_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            constexpr int CPW = BN / (Threads<BN, Epi>::NEPI / 4);   // accumulator columns copied by each warp of a lane quarter
            const int cbeg = (e >> 2) * CPW;
#pragma unroll
            for (int c0 = cbeg; c0 < cbeg + CPW; c0 += 32) {
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                for (int a = 1; a < nacc; ++a) {
                    float w[32];
                    tmem_ld32(trow + (uint32_t)(a * BN + c0), w);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += w[j];
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(Cs + row * C::LDC + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (probe && threadIdx.x == 64) probe[5] = gtimer();
    }
    if constexpr (KS > 1) {
        if (warp < 2) {   // the producer / MMA warps take part in the cluster barrier too (it counts every thread)
            asm volatile("barrier.cluster.arrive.release;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        }
    }
    __syncthreads();
    if constexpr (!Epi::kDirect) Epi::template apply<C>(ep, Cs, m0, n0);
    if constexpr (MC) cluster_sync_all();   // peers may still signal this CTA's barriers: nobody leaves before everybody is done
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side: tensor maps ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void* ptr; int rows, cols, ld, box_rows;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h ^= std::hash<long long>()(((long long)k.rows << 32) ^ k.cols) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        h ^= std::hash<long long>()(((long long)k.ld << 32) ^ k.box_rows) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        return h;
    }
};
typedef std::unordered_map<MapKey, CUtensorMap, MapKeyHash> MapCache;

// bf16 row-major [rows, cols] with row stride ld elements; box = 64 columns x box_rows rows, 128-byte swizzle, zero fill.
inline const CUtensorMap* get_map(MapCache& cache, const void* ptr, int rows, int cols, int ld, int box_rows) {
    MapKey key = {ptr, rows, cols, ld, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return nullptr;
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nullptr;
    return &cache.emplace(key, m).first->second;
}

// The same matrix seen as [K / 64 blocks][rows][64 columns]: one box {64, box_rows, kblocks} lands `kblocks` consecutive K-block tiles
// (each box_rows x 128 bytes, 128-byte swizzle) in shared memory with ONE TMA instruction.  Key: box_rows carries the block count too.
inline const CUtensorMap* get_map3d(MapCache& cache, const void* ptr, int rows, int cols, int ld, int box_rows, int kblocks) {
    MapKey key = {ptr, rows, cols, ld, -(box_rows * 64 + kblocks)};
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return nullptr;
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)BK, (cuuint64_t)rows, (cuuint64_t)(cols / BK)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)BK * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)kblocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nullptr;
    return &cache.emplace(key, m).first->second;
}

// MN == false: A [M, K], B [N, K] row-major (K-major operands).  MN == true: A = X [K rows, M features], B = Y [K rows, N features]
// row-major; C[M, N] = X^T . Y.
template <int BN, class Epi, int KS = 1, int CX = 1, int CY = 1, bool MN = false>
inline cudaError_t launch(MapCache& cache, cudaStream_t st, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K,
                          const typename Epi::Params& ep, bool pdl = false, uint32_t fmt = 0) {
    if (M <= 0) return cudaSuccess;
    constexpr int NT = Threads<BN, Epi>::N;
    using C = Cfg<BN, NT>;
    constexpr int SMEM = C::SMEM_BYTES + (KS > 1 ? KS * 32 * BN * 4 : 0);
    if (cache.size() > 32768) cache.clear();   // before either lookup: element pointers stay valid across inserts, not across clear
    const int a_rows = (CX == 1 && M <= 64) ? ((M + 7) & ~7) : BM;
#ifdef S2VT_TMA_SPLIT
    const CUtensorMap* ma = get_map(cache, A, M, K, lda, a_rows < 64 ? a_rows : 64);
    const CUtensorMap* mb = get_map(cache, B, N, K, ldb, BN < 64 ? BN : 64);
#else
    const CUtensorMap* ma = MN ? get_map(cache, A, K, M, lda, 64) : get_map(cache, A, M, K, lda, CX == 1 ? a_rows : BM / CX);
    const CUtensorMap* mb = MN ? get_map(cache, B, K, N, ldb, 64) : get_map(cache, B, N, K, ldb, BN / CY);
#endif
    if (!ma || !mb) return cudaErrorInvalidValue;
    auto kern = gemm_tc_kernel<BN, Epi, KS, CX, CY, MN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(N / BN, ((M + BM - 1) / BM + CY - 1) / CY * CY, KS);   // whole clusters; surplus row tiles are all-padding
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (KS > 1 || CX * CY > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CX; attr[na].val.clusterDim.y = CY; attr[na].val.clusterDim.z = KS;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, *ma, *mb, K, a_rows, fmt, ep);
}

}  // namespace tc
