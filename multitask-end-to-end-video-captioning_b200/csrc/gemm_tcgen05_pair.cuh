// CTA-pair (cta_group::2) tcgen05 GEMM for the large batched products:  C[M,N] = A[M,K] . B[N,K]^T  (or C = X^T . Y, MN-major).
//
// The single-CTA kernel of gemm_tcgen05.cuh is bound by the L2 -> SM fill (~42 B/clk/SM): a 128 x 256 tile needs 48 KB per K-block
// for 4.2 MFLOP (85 flop per fetched byte).  Here two CTAs on the two SMs of a TPC (cluster dims (2,1,1)) compute one 256 x 256 tile
// with ONE tcgen05.mma.cta_group::2 (M = 256, N = 256) per K step: CTA r of the pair holds rows [128 r, 128 r + 128) of A and rows
// [128 r, 128 r + 128) of the B tile, the tensor cores read both halves of B, each CTA's TMEM receives its own 128 x 256 part of the
// accumulator.  Per CTA and K-block that is 32 KB for the same 4.2 MFLOP -- 131 flop per fetched byte.
//
// Protocol (CUTLASS / DeepGEMM 2-SM scheme): both CTAs run a TMA producer warp; every load completes on the LEADER's (rank 0) `full`
// barrier (cp.async.bulk.tensor...cta_group::2, barrier address mapped to rank 0), which expects the bytes of both CTAs; only the
// leader's MMA warp issues tcgen05.mma, and its tcgen05.commit is multicast to the `empty` / `tmem_full` barriers of both CTAs.
// Epilogue: as in gemm_tc_kernel (TMEM -> fp32 smem tile -> staged epilogue functor), each CTA on its own 128 rows.
#pragma once
#include "gemm_tcgen05.cuh"

namespace tc {

constexpr int PAIR_BN = 256, PAIR_STAGES = 6, PAIR_STAGE_BYTES = 2 * 128 * 128;   // per CTA: A 128 x 64 + B half 128 x 64, bf16 / fp16

__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <class Epi, bool MN>
__global__ void __launch_bounds__(Threads<PAIR_BN, Epi>::N) gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                                int K, uint32_t fmt, typename Epi::Params ep) {
    static_assert(!Epi::kDirect, "staged epilogues only");
    using C = Cfg<PAIR_BN, Threads<PAIR_BN, Epi>::N>;
    constexpr int A_BYTES = 128 * 128, STAGE = PAIR_STAGE_BYTES;
    // M = 256 (both CTAs), N = 256, F32 accumulate; MN-major sets the two transpose bits
    const uint32_t IDESC = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24) | (MN ? ((1u << 15) | (1u << 16)) : 0u)) & ~fmt;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int PIPE = PAIR_STAGES * STAGE;
    constexpr int MAIN = PIPE > C::EPI_BYTES ? PIPE : C::EPI_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + MAIN);
    uint64_t* empty = full + PAIR_STAGES;
    uint64_t* tmem_full = empty + PAIR_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* Cs = reinterpret_cast<float*>(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * PAIR_BN;          // blockIdx.x = 2 * pair + rank: cluster dims (2, 1, 1)
    const int nb = n0 + 128 * (int)rank;                                  // this CTA's half of the B tile
    // gridDim.z > 1: split-K -- slice z multiplies K-blocks [kb0, kb0 + KBL); the epilogue functor must then accumulate atomically
    const int KBT = MN ? (K + BK - 1) / BK : K / BK, KBS = (KBT + (int)gridDim.z - 1) / (int)gridDim.z;
    const int kb0 = (int)blockIdx.z * KBS;
    const int KBL = KBT - kb0 < KBS ? (KBT - kb0 > 0 ? KBT - kb0 : 0) : KBS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < PAIR_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // both CTAs of the pair allocate (same columns in both)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(PAIR_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();                                                   // the peer's barriers are initialised before anything signals them
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        const bool leader_lane = elect_one();
        for (int i = 0; i < KBL; ++i) {
            const int s = i % PAIR_STAGES;
            if (i >= PAIR_STAGES) mbar_wait(empty + s, ((i / PAIR_STAGES) - 1) & 1);
            if (leader_lane) {
                const uint32_t lbar = cluster_map(smem_u32(full + s), 0);                      // the leader CTA's barrier of this stage
                if (rank == 0) mbar_expect_tx(full + s, 2u * (uint32_t)STAGE);                 // bytes of BOTH CTAs
                unsigned char* a = smem + s * STAGE;
                unsigned char* b = a + A_BYTES;
                if constexpr (MN) {
                    tma_load_2d_pair(a, &mapA, lbar, m0, (kb0 + i) * BK);
                    tma_load_2d_pair(a + 8192, &mapA, lbar, m0 + 64, (kb0 + i) * BK);
                    tma_load_2d_pair(b, &mapB, lbar, nb, (kb0 + i) * BK);
                    tma_load_2d_pair(b + 8192, &mapB, lbar, nb + 64, (kb0 + i) * BK);
                } else {
                    tma_load_2d_pair(a, &mapA, lbar, (kb0 + i) * BK, m0);
                    tma_load_2d_pair(b, &mapB, lbar, (kb0 + i) * BK, nb);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0) {
            const bool leader_lane = elect_one();
            for (int i = 0; i < KBL; ++i) {
                const int s = i % PAIR_STAGES;
                mbar_wait(full + s, (i / PAIR_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = smem_u32(smem + s * STAGE);
                const uint64_t adesc = MN ? make_desc_mn(a) : make_desc(a), bdesc = MN ? make_desc_mn(a + A_BYTES) : make_desc(a + A_BYTES);
                constexpr int KADV = MN ? 128 : 2;
                if (leader_lane) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) mma_pair(tmem_base, adesc + KADV * k, bdesc + KADV * k, IDESC, i > 0 || k != 0);
                    mma_commit_pair(empty + s);
                }
            }
            if (leader_lane && KBL > 0) mma_commit_pair(tmem_full);
            __syncwarp();
        }
    } else {
        const int e = warp - 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        if (KBL > 0) mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        constexpr int CPW = PAIR_BN / (Threads<PAIR_BN, Epi>::NEPI / 4);
        const int cbeg = (e >> 2) * CPW;
#pragma unroll
        for (int c0 = cbeg; c0 < cbeg + CPW; c0 += 32) {
            float v[32];
            tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(Cs + row * C::LDC + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (KBL > 0) Epi::template apply<C>(ep, Cs, m0, n0);
    cluster_sync_all();                                                   // nobody leaves (or frees TMEM) while the peer may still signal / be read
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(PAIR_BN) : "memory");
    }
}

// N % 256 == 0.  MN == false: A [M, K], B [N, K] row-major; MN == true: A = X [K rows, M features], B = Y [K rows, N features].
template <class Epi, bool MN = false>
inline cudaError_t launch_pair(MapCache& cache, cudaStream_t st, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K,
                               const typename Epi::Params& ep, bool pdl, uint32_t fmt, int ksplit = 1) {
    if (M <= 0) return cudaSuccess;
    constexpr int NT = Threads<PAIR_BN, Epi>::N;
    using C = Cfg<PAIR_BN, NT>;
    constexpr int PIPE = PAIR_STAGES * PAIR_STAGE_BYTES;
    constexpr int SMEM = (PIPE > C::EPI_BYTES ? PIPE : C::EPI_BYTES) + 256 + 1024;
    if (cache.size() > 32768) cache.clear();
    const CUtensorMap* ma = MN ? get_map(cache, A, K, M, lda, 64) : get_map(cache, A, M, K, lda, 128);
    const CUtensorMap* mb = MN ? get_map(cache, B, K, N, ldb, 64) : get_map(cache, B, N, K, ldb, 128);
    if (!ma || !mb) return cudaErrorInvalidValue;
    auto kern = gemm_tc_pair_kernel<Epi, MN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((M + 255) / 256) * 2, N / PAIR_BN, ksplit < 1 ? 1 : ksplit);
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
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    static const bool debug = getenv("S2VT_DEBUG_CHAIN") != nullptr;
    if (debug) {
        int ncl = -1;
        cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
        fprintf(stderr, "s2vt pair gemm: M=%d N=%d K=%d grid=(%u,%u) threads=%d smem=%d max active clusters=%d (%s)\n", M, N, K, cfg.gridDim.x, cfg.gridDim.y, NT, SMEM, ncl,
                cudaGetErrorString(qe));
        (void)cudaGetLastError();
    }
    return cudaLaunchKernelEx(&cfg, kern, *ma, *mb, K, fmt, ep);
}

}  // namespace tc
