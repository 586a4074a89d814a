// Weights-stationary, software-pipelined recurrent chain for MORE than 128 rows (the 320-row teacher-forced LSTM2 chain of a
// REINFORCE iteration):   for s in 0 .. nsteps-1 :  C_s = A_s . B^T  ->  fused BasicLSTMCell epilogue,   A_s = rows of step s-1's output.
//
// What bounds the plain persistent chain (gemm_tcgen05_chain.cuh) at 320 rows is, per step: a grid barrier (1.4 us), the first-byte
// latency after it (1.4 us), 512 KB of operand fill per CTA (5.4 us: half of it the SAME weights every step) and a 4.4 us cell
// epilogue -- one after the other.  This kernel changes three things:
//   * the CTA's weight slab (96 gate columns x K, 192 KB) is loaded ONCE and stays in shared memory: a step streams activations only;
//     grid = ceil(N / 96) column slabs x 3 row groups (43 x 3 = 129 CTAs at H = 1000, one per SM);
//   * every row group is cut into two independent HALVES (<= 64 rows, one tcgen05.mma M = 64 tile and one TMEM accumulator each).
//     Step s+1 of a half needs step s of THAT half only, so while the epilogue warps run the cell of half 0 the TMA / MMA warps
//     already fetch and multiply half 1, and vice versa: barrier latency, fill and MMAs hide under the other half's epilogue;
//   * the dependency is a counter per (row group, half) that only the 43 CTAs of that row group touch -- arrivals from the epilogue
//     warps (red.release.gpu), polled by the producer warp alone (ld.acquire.gpu + fence.proxy.async before the TMA reads); no
//     CTA-wide or grid-wide barrier exists inside the time loop.
// Accumulation order equals the other tcgen05 kernels (one accumulator, K ascending), so results are bit-identical to them.
//
// tcgen05.mma M = 64 (cta_group::1) places accumulator row m in TMEM lane  (m % 16) + 32 (m / 16)  (cute tmem_frg_1sm, "half
// sub-partition" atom): warp quarter q holds rows 16 q .. 16 q + 15 in its lanes 0 .. 15.
#pragma once
#include "gemm_tcgen05_chain.cuh"

namespace tc {

constexpr int WS2_KB_MAX = 16, WS2_A_STAGE = 64 * 128;
// BN: gate columns of the resident slab (96: 43 slabs x 3 row groups; 64: 64 slabs x 2 row groups, a smaller slab and a deeper ring)
template <int BN> struct Ws2Cfg {
    static constexpr int STAGES = (227 * 1024 - 2048 - WS2_KB_MAX * BN * 128) / WS2_A_STAGE;      // what the slab leaves of the 227 KB
    static constexpr int EPI_WARPS = 4 * (BN / 32), THREADS = 64 + 32 * EPI_WARPS, W_TILE = BN * 128;
    static constexpr int SMEM = STAGES * WS2_A_STAGE + WS2_KB_MAX * W_TILE + 256 + 1024;
    static constexpr int RG = BN >= 96 ? 3 : 2;
};

#ifdef S2VT_CHAIN_PROBE
#define WS2_PROBE(stmt) do { stmt; } while (0)
#else
#define WS2_PROBE(stmt) do { } while (0)
#endif

template <class Epi, int WS2_BN = 96>
__global__ void __launch_bounds__(Ws2Cfg<WS2_BN>::THREADS) gemm_tc_ws2_chain_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                        int K, int M, int rpg, int n_limit, int a_row0, int a_row_stride,
                                                                        const typename Epi::Params* __restrict__ steps, int nsteps,
                                                                        unsigned* __restrict__ flags, uint32_t fmt) {
    static_assert(Epi::kDirect, "register epilogue");
    constexpr int WS2_STAGES = Ws2Cfg<WS2_BN>::STAGES, WS2_EPI_WARPS = Ws2Cfg<WS2_BN>::EPI_WARPS, WS2_W_TILE = Ws2Cfg<WS2_BN>::W_TILE;
    // instruction descriptor: F32 accumulate, BF16 formats (cleared to F16 by fmt), K-major operands, N = 96, M = 64
    const uint32_t IDESC = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(WS2_BN >> 3) << 17) | ((uint32_t)(64 >> 4) << 24)) & ~fmt;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;                                           // [STAGES][64 rows x 128 B]   activation K-blocks
    unsigned char* wsm = smem + WS2_STAGES * WS2_A_STAGE;                 // [KBL][96 rows x 128 B]      the resident weight slab
    uint64_t* full = reinterpret_cast<uint64_t*>(wsm + WS2_KB_MAX * WS2_W_TILE);
    uint64_t* empty = full + WS2_STAGES;
    uint64_t* wfull = empty + WS2_STAGES;
    uint64_t* acc_full = wfull + 1;                                       // [2] MMAs of a half retired -> epilogue
    uint64_t* acc_free = acc_full + 2;                                    // [2] epilogue has read the accumulator -> next step's MMAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WS2_BN, rg = blockIdx.y;
    const int KBL = K / BK;
    const unsigned ncol = gridDim.x;
    const int hr = rpg >> 1;                                              // rows of a half (box height of mapA)
    int row_lo[2], valid[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        row_lo[hf] = rg * rpg + hf * hr;
        int v = M - row_lo[hf];
        valid[hf] = v < 0 ? 0 : (v > hr ? hr : v);
    }
#ifdef S2VT_CHAIN_PROBE
    __shared__ unsigned long long* probe;
    if (threadIdx.x == 0) {
        probe = nullptr;
        if (g_probe && blockIdx.x == 0 && blockIdx.y == 0) {
            unsigned long long slot = atomicAdd(g_probe, (unsigned long long)nsteps);
            if (slot + nsteps < 4000) probe = g_probe + 8 * (slot + 1);
        }
    }
#endif

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < WS2_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(wfull, 1);
        for (int hf = 0; hf < 2; ++hf) { mbar_init(acc_full + hf, 1); mbar_init(acc_free + hf, WS2_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the weight slab, once: weights never depend on the previous kernel (only s2vt_refresh writes them)
        mbar_expect_tx(wfull, (uint32_t)(KBL * WS2_W_TILE));
        for (int i = 0; i < KBL; ++i) tma_load_2d_raw(wsm + i * WS2_W_TILE, &mapB, wfull, i * BK, n0);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");                    // everything below may read what the previous kernel wrote

    if (warp == 0) {
        // ---- producer: activation K-blocks of (step, half), as soon as that half's previous step is published by its row group
        const bool leader = elect_one();
        int g = 0;                                                        // running K-block counter (ring phases continue across steps)
        for (int s = 0; s < nsteps; ++s) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (valid[hf] == 0) continue;
                if (s > 0) {
                    if (lane == 0) {
                        const unsigned* f = flags + (rg * 2 + hf);
                        const unsigned target = (unsigned)s * ncol;
                        long long t0 = clock64();
                        for (unsigned spins = 1; ld_acquire_gpu(f) < target; ++spins) {
                            if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
                                printf("s2vt: ws2 chain flag timed out (step %d half %d block %d,%d)\n", s, hf, blockIdx.x, blockIdx.y);
                                __trap();
                            }
                        }
                    }
                    __syncwarp();
                }
                WS2_PROBE(if (probe && lane == 0) probe[8 * s + 4 * hf + 0] = gtimer());      // dependency satisfied, loads start
                if (leader) asm volatile("fence.proxy.async;" ::: "memory");   // rows written through the generic proxy by other SMs are read by TMA
                const int arow = a_row0 + s * a_row_stride + row_lo[hf];
                for (int i = 0; i < KBL; ++i, ++g) {
                    const int st = g % WS2_STAGES;
                    if (g >= WS2_STAGES) mbar_wait(empty + st, ((g / WS2_STAGES) - 1) & 1);
                    if (leader) {
                        mbar_expect_tx(full + st, (uint32_t)(hr * 128));
                        tma_load_2d_raw(ring + st * WS2_A_STAGE, &mapA, full + st, i * BK, arow);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issuer: one M = 64 tile per half, accumulators at TMEM columns 0 and 128
        const bool leader = elect_one();
        mbar_wait(wfull, 0);
        const uint64_t adesc0 = make_desc(smem_u32(ring)), bdesc0 = make_desc(smem_u32(wsm));   // start-address field: +bytes/16 per stage / K-block (no carry: smem < 256 KB)
        int g = 0;
        for (int s = 0; s < nsteps; ++s) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (valid[hf] == 0) continue;
                if (s > 0) mbar_wait(acc_free + hf, (s - 1) & 1);         // the epilogue of step s-1 has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int i = 0; i < KBL; ++i, ++g) {
                    const int st = g % WS2_STAGES;
                    mbar_wait(full + st, (g / WS2_STAGES) & 1);      // (TMA completion: no tcgen05 fence needed, the per-half fence above orders the TMEM reuse)
                    const uint64_t adesc = adesc0 + (uint64_t)((st * WS2_A_STAGE) >> 4), bdesc = bdesc0 + (uint64_t)((i * WS2_W_TILE) >> 4);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base + (uint32_t)(hf * 128), adesc + 2 * k, bdesc + 2 * k, IDESC, i > 0 || k != 0);
                        mma_commit(empty + st);
                    }
                }
                if (leader) {
                    mma_commit(acc_full + hf);
                    WS2_PROBE(if (probe) probe[8 * s + 4 * hf + 1] = gtimer());                  // all MMAs of the half issued
                }
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: 12 warps = 4 TMEM lane quarters x 3 chunks of 32 gate columns (8 units); lanes 0..15 of a warp own a row each
        const int e = warp - 2, q = warp & 3, chunk = e >> 2;
        const int gc = n0 + 32 * chunk;
        const bool col_ok = gc < n_limit;
        for (int s = 0; s < nsteps; ++s) {
            const typename Epi::Params& ep = steps[s];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (valid[hf] == 0) continue;
                const int rih = 16 * q + lane;                            // row inside the half (lanes >= 16 hold no accumulator row)
                const int gr = (lane < 16 && rih < valid[hf] && col_ok) ? row_lo[hf] + rih : M;   // M: "no row" for the epilogue functor
                typename Epi::Pre pre;
                Epi::prefetch(ep, gr, gc, pre);
                mbar_wait(acc_full + hf, s & 1);
                WS2_PROBE(if (probe && threadIdx.x == 64) probe[8 * s + 4 * hf + 2] = gtimer());        // accumulator ready, epilogue starts
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(hf * 128 + 32 * chunk), v);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_free + hf)) : "memory");
                Epi::direct(ep, gr, gc, v, pre);
                // publish this half of step s to the row group: every epilogue thread's stores, then ONE release arrival per CTA
                asm volatile("bar.sync 1, %0;" ::"n"(32 * WS2_EPI_WARPS) : "memory");
                if (threadIdx.x == 64) {
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(flags + (rg * 2 + hf)), "r"(1u) : "memory");
                    WS2_PROBE(if (probe) probe[8 * s + 4 * hf + 3] = gtimer());                      // half published
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// Host launcher.  A: one buffer of `a_total_rows` rows; step s reads rows [a_row0 + s * a_row_stride, + M).  B: [N rows, K] K-major.
// n_logical: gate columns that carry weights (4 H); columns beyond ceil-to-96 of it are never computed, so the caller must not rely
// on them (the engine's padded units between 4 H and N produce zeros in the plain kernels; here they are covered up to N because
// the slab count is taken from N).  Returns cudaErrorLaunchOutOfResources (nothing launched) when the shape does not fit: M > 384,
// K > 1024, or the grid cannot be co-resident.
template <class Epi, int WS2_BN = 96>
inline cudaError_t launch_ws2_chain(MapCache& cache, cudaStream_t st, const bf16* A, int lda, int a_total_rows, int a_row0, int a_row_stride, const bf16* B,
                                    int ldb, int M, int N, int K, const typename Epi::Params* steps_dev, int nsteps, unsigned* flags, bool pdl, uint32_t fmt) {
    constexpr int RG = Ws2Cfg<WS2_BN>::RG, WS2_SMEM = Ws2Cfg<WS2_BN>::SMEM, WS2_THREADS = Ws2Cfg<WS2_BN>::THREADS;
    if (K % BK != 0 || K / BK > WS2_KB_MAX || M <= 128 || M > RG * 128) return cudaErrorLaunchOutOfResources;
    const int rpg = ((M + RG - 1) / RG + 15) & ~15;                       // rows per row group, halves are multiples of 8 rows
    const int ncol = (N + WS2_BN - 1) / WS2_BN;
    if (cache.size() > 32768) cache.clear();
    const CUtensorMap* ma = get_map(cache, A, a_total_rows, K, lda, rpg / 2);
    const CUtensorMap* mb = get_map(cache, B, N, K, ldb, WS2_BN);
    if (!ma || !mb) return cudaErrorInvalidValue;
    auto kern = gemm_tc_ws2_chain_kernel<Epi, WS2_BN>;
    static int max_ctas = -1;
    if (max_ctas < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WS2_SMEM);
        if (e != cudaSuccess) return e;
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WS2_THREADS, WS2_SMEM);
        if (e != cudaSuccess) return e;
        max_ctas = per_sm * sms;
    }
    static const bool debug = getenv("S2VT_DEBUG_CHAIN") != nullptr;
    if (debug) fprintf(stderr, "s2vt ws2 chain: rows=%d rpg=%d N=%d K=%d steps=%d grid=%d x %d co-resident limit=%d\n", M, rpg, N, K, nsteps, ncol, RG, max_ctas);
    if (ncol * RG > max_ctas) return cudaErrorLaunchOutOfResources;
    cudaError_t e = cudaMemsetAsync(flags, 0, 2 * RG * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncol, RG, 1);
    cfg.blockDim = dim3(WS2_THREADS);
    cfg.dynamicSmemBytes = WS2_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, *ma, *mb, K, M, rpg, N, a_row0, a_row_stride, steps_dev, nsteps, flags, fmt);
}

// -------------------------------------------------------------------------------------------------------------------------------
// The same scheme for the BPTT chain:  dh_rec[rows, units] = dG(t+1)[rows, 4 H] . W_h^T  ->  fused BasicLSTMCell backward.
// The contraction is 4 H = 4096 long, so it is split over KS = 4 CTAs (K slices of 1024 gate columns): CTA (c, rg, ks) keeps the
// 96-unit x 1024 slab of W_h in shared memory (192 KB), streams the K slice of its half's gate-gradient rows, and leaves a partial
// 64 x 96 tile in TMEM.  Instead of the 4-CTA DSMEM exchange of the ring chain (cluster barrier, 7 us epilogue) the partial tiles go
// through L2: every CTA stores its partial into its own slot of a scratch matrix (plain vector stores, deterministic), announces it on
// a counter per (row group, half, column slab), and -- once the four partials are there -- finishes a quarter of the half's rows:
// sums the four slots in slice order and runs the cell backward.  The wait for the three peers costs an L2 round trip, but it sits in
// one half's dependency cycle while the TMA / MMA warps are already busy with the other half.
// grid (ceil(N / 96), 3, 4) = 11 x 3 x 4 = 132 CTAs at H = 1000; flags: [2 * 3] step flags + [2 * 3 * ncol] partial counters.
template <class Epi>
__global__ void __launch_bounds__(Ws2Cfg<96>::THREADS) gemm_tc_ws2_bwd_chain_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                                    int K, int M, int rpg, int n_limit, int a_row0, int a_row_stride,
                                                                                    const typename Epi::Params* __restrict__ steps, int nsteps,
                                                                                    unsigned* __restrict__ flags, float* __restrict__ scratch, uint32_t fmt) {
    constexpr int WS2_BN = 96, KS = 4;
    constexpr int WS2_STAGES = Ws2Cfg<WS2_BN>::STAGES, WS2_EPI_WARPS = Ws2Cfg<WS2_BN>::EPI_WARPS, WS2_W_TILE = Ws2Cfg<WS2_BN>::W_TILE;
    static_assert(Epi::kDirect && Epi::kUnitsPerChunk == 8, "cell-backward register epilogue (8 units per call)");
    const uint32_t IDESC = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(WS2_BN >> 3) << 17) | ((uint32_t)(64 >> 4) << 24)) & ~fmt;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = smem;
    unsigned char* wsm = smem + WS2_STAGES * WS2_A_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(wsm + WS2_KB_MAX * WS2_W_TILE);
    uint64_t* empty = full + WS2_STAGES;
    uint64_t* wfull = empty + WS2_STAGES;
    uint64_t* acc_full = wfull + 1;
    uint64_t* acc_free = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WS2_BN, rg = blockIdx.y, ks = blockIdx.z;
    const int KBL = K / BK / KS, kb0 = ks * KBL;                          // this CTA's K-blocks
    const unsigned ncol = gridDim.x;
    const int hr = rpg >> 1;
    const int ldn = (int)ncol * WS2_BN;                                    // scratch row length (units, padded to whole slabs)
    int row_lo[2], valid[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        row_lo[hf] = rg * rpg + hf * hr;
        int v = M - row_lo[hf];
        valid[hf] = v < 0 ? 0 : (v > hr ? hr : v);
    }
    unsigned* step_flag = flags;                                           // [rg * 2 + hf]: CTAs of the row group that finished the step of this half
    unsigned* part_cnt = flags + 8;                                        // [(rg * 2 + hf) * ncol + c]: partial tiles stored for this column slab

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < WS2_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(wfull, 1);
        for (int hf = 0; hf < 2; ++hf) { mbar_init(acc_full + hf, 1); mbar_init(acc_free + hf, WS2_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(wfull, (uint32_t)(KBL * WS2_W_TILE));
        for (int i = 0; i < KBL; ++i) tma_load_2d_raw(wsm + i * WS2_W_TILE, &mapB, wfull, (kb0 + i) * BK, n0);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        const bool leader = elect_one();
        int g = 0;
        for (int s = 0; s < nsteps; ++s) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (valid[hf] == 0) continue;
                if (s > 0) {
                    if (lane == 0) {
                        const unsigned* f = step_flag + (rg * 2 + hf);
                        const unsigned target = (unsigned)s * ncol * KS;
                        long long t0 = clock64();
                        for (unsigned spins = 1; ld_acquire_gpu(f) < target; ++spins) {
                            if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
                                printf("s2vt: ws2 bwd chain flag timed out (step %d half %d block %d,%d,%d)\n", s, hf, blockIdx.x, blockIdx.y, blockIdx.z);
                                __trap();
                            }
                        }
                    }
                    __syncwarp();
                }
                if (leader) asm volatile("fence.proxy.async;" ::: "memory");
                const int arow = a_row0 + s * a_row_stride + row_lo[hf];
                for (int i = 0; i < KBL; ++i, ++g) {
                    const int st = g % WS2_STAGES;
                    if (g >= WS2_STAGES) mbar_wait(empty + st, ((g / WS2_STAGES) - 1) & 1);
                    if (leader) {
                        mbar_expect_tx(full + st, (uint32_t)(hr * 128));
                        tma_load_2d_raw(ring + st * WS2_A_STAGE, &mapA, full + st, (kb0 + i) * BK, arow);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const bool leader = elect_one();
        mbar_wait(wfull, 0);
        const uint64_t adesc0 = make_desc(smem_u32(ring)), bdesc0 = make_desc(smem_u32(wsm));
        int g = 0;
        for (int s = 0; s < nsteps; ++s) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (valid[hf] == 0) continue;
                if (s > 0) mbar_wait(acc_free + hf, (s - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int i = 0; i < KBL; ++i, ++g) {
                    const int st = g % WS2_STAGES;
                    mbar_wait(full + st, (g / WS2_STAGES) & 1);
                    const uint64_t adesc = adesc0 + (uint64_t)((st * WS2_A_STAGE) >> 4), bdesc = bdesc0 + (uint64_t)((i * WS2_W_TILE) >> 4);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base + (uint32_t)(hf * 128), adesc + 2 * k, bdesc + 2 * k, IDESC, i > 0 || k != 0);
                        mma_commit(empty + st);
                    }
                }
                if (leader) mma_commit(acc_full + hf);
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2, q = warp & 3, chunk = e >> 2;
        const int t = (int)threadIdx.x - 64;                              // 0 .. 383: finishing task id
        for (int s = 0; s < nsteps; ++s) {
            const typename Epi::Params& ep = steps[s];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (valid[hf] == 0) continue;
                // this CTA finishes rows [ks * rq, ks * rq + rq) of the half, 12 tasks of 8 units per row
                const int rq = (valid[hf] + KS - 1) / KS;
                const int frow = ks * rq + t / 12, fu = n0 + 8 * (t % 12);
                const bool ftask = t / 12 < rq && frow < valid[hf] && fu < n_limit;
                const int fgr = ftask ? row_lo[hf] + frow : M;
                typename Epi::Pre pre;
                Epi::prefetch(ep, fgr, fu, pre);                          // gates, c, dc, dh_ext of the task: in flight while the MMAs run
                mbar_wait(acc_full + hf, s & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(hf * 128 + 32 * chunk), v);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_free + hf)) : "memory");
                // ---- partial tile -> this K slice's slot of the scratch matrix [rg][hf][ks][64 rows][ldn]
                float* slot = scratch + ((size_t)((rg * 2 + hf) * KS + ks) * 64) * ldn;
                const int rih = 16 * q + lane;
                if (lane < 16 && rih < valid[hf]) {
                    float4* d = reinterpret_cast<float4*>(slot + (size_t)rih * ldn + n0 + 32 * chunk);
#pragma unroll
                    for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * WS2_EPI_WARPS) : "memory");
                unsigned* pc = part_cnt + ((rg * 2 + hf) * ncol + blockIdx.x);
                if (threadIdx.x == 64) {
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(pc), "r"(1u) : "memory");
                    const unsigned target = (unsigned)(s + 1) * KS;
                    long long t0 = clock64();
                    for (unsigned spins = 1; ld_acquire_gpu(pc) < target; ++spins) {
                        if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
                            printf("s2vt: ws2 bwd chain partial counter timed out (step %d half %d block %d,%d,%d)\n", s, hf, blockIdx.x, blockIdx.y, blockIdx.z);
                            __trap();
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * WS2_EPI_WARPS) : "memory");      // the four partials are visible (acquire by thread 64 + barrier)
                // ---- finish: sum the four slices in order, cell backward for 8 units of one row
                if (ftask) {
                    const float* base = scratch + ((size_t)((rg * 2 + hf) * KS) * 64 + frow) * ldn + fu;
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k4 = 0; k4 < KS; ++k4) {
                        const float4* pp = reinterpret_cast<const float4*>(base + (size_t)k4 * 64 * ldn);
                        float4 x0, x1;
                        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "l"(pp));
                        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "l"(pp + 1));
                        acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w; acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
                    }
                    Epi::direct(ep, fgr, fu, acc, pre);
                }
                // publish the finished quarter to the row group
                asm volatile("bar.sync 1, %0;" ::"n"(32 * WS2_EPI_WARPS) : "memory");
                if (threadIdx.x == 64) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(step_flag + (rg * 2 + hf)), "r"(1u) : "memory");
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// floats of scratch the backward chain needs for M rows and N units
inline size_t ws2_bwd_scratch_floats(int N) { return (size_t)3 * 2 * 4 * 64 * (size_t)((N + 95) / 96 * 96); }

// B: [N units, K] K-major (K = 4 H gate columns).  flags: >= 8 + 6 * ceil(N / 96) unsigned.
template <class Epi>
inline cudaError_t launch_ws2_bwd_chain(MapCache& cache, cudaStream_t st, const bf16* A, int lda, int a_total_rows, int a_row0, int a_row_stride, const bf16* B,
                                        int ldb, int M, int N, int K, const typename Epi::Params* steps_dev, int nsteps, unsigned* flags, int flags_cap,
                                        float* scratch, bool pdl, uint32_t fmt) {
    constexpr int RG = 3, KS = 4, BN = 96;
    if (K % (BK * KS) != 0 || K / BK / KS > WS2_KB_MAX || M <= 128 || M > RG * 128 || !scratch) return cudaErrorLaunchOutOfResources;
    const int rpg = ((M + RG - 1) / RG + 15) & ~15;
    const int ncol = (N + BN - 1) / BN;
    if (8 + 2 * RG * ncol > flags_cap) return cudaErrorLaunchOutOfResources;
    if (cache.size() > 32768) cache.clear();
    const CUtensorMap* ma = get_map(cache, A, a_total_rows, K, lda, rpg / 2);
    const CUtensorMap* mb = get_map(cache, B, N, K, ldb, BN);
    if (!ma || !mb) return cudaErrorInvalidValue;
    auto kern = gemm_tc_ws2_bwd_chain_kernel<Epi>;
    constexpr int SMEM = Ws2Cfg<BN>::SMEM, NT = Ws2Cfg<BN>::THREADS;
    static int max_ctas = -1;
    if (max_ctas < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, SMEM);
        if (e != cudaSuccess) return e;
        max_ctas = per_sm * sms;
    }
    static const bool debug = getenv("S2VT_DEBUG_CHAIN") != nullptr;
    if (debug) fprintf(stderr, "s2vt ws2 bwd chain: rows=%d rpg=%d N=%d K=%d steps=%d grid=%d x %d x %d co-resident limit=%d\n", M, rpg, N, K, nsteps, ncol, RG, KS, max_ctas);
    if (ncol * RG * KS > max_ctas) return cudaErrorLaunchOutOfResources;
    cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)(8 + 2 * RG * ncol) * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncol, RG, KS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, *ma, *mb, K, M, rpg, N, a_row0, a_row_stride, steps_dev, nsteps, flags, scratch, fmt);
}

}  // namespace tc
