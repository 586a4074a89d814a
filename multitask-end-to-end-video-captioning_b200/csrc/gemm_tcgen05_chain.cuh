// Persistent recurrent-step kernel: ONE launch walks a whole chain of dependent time steps
//     for s in 0 .. nsteps-1 :   C_s = A_s . B^T  ->  fused LSTM-cell epilogue (forward or backward)
// where A_s (the previous step's hidden state / gate gradients) lives at row  a_row0 + s * a_row_stride  of one buffer
// and B (the recurrent weights) is the same for every step.  The CTAs are the tiles of gemm_tc_kernel (same TMA /
// tcgen05 / TMEM pipeline, same register epilogues, optional split-K cluster of 4) and stay resident; consecutive steps
// are separated by a grid-wide barrier (atomic counter in global memory) instead of a kernel boundary.  Barriers,
// TMEM allocation and tensor-map fetches are paid once per chain; the weight tiles of the next step are requested
// before the barrier (they never depend on it).
//
// Memory ordering across the barrier: epilogue stores (generic proxy) -> __syncthreads -> thread 0: __threadfence +
// atomicAdd ... other CTAs: acquire spin -> fence.proxy.async -> TMA loads (async proxy) of the rows just written.
// Requires every CTA of the grid to be co-resident (checked by the host with the occupancy API).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "gemm_tcgen05.cuh"

namespace tc {

// Per-step phase timestamps (scripts/probe_chains.py) are compiled in only with -DS2VT_CHAIN_PROBE: the production kernel carries
// no probe code.
#ifdef S2VT_CHAIN_PROBE
#define CHAIN_PROBE(stmt) do { stmt; } while (0)
#else
#define CHAIN_PROBE(stmt) do { } while (0)
#endif

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Watcher for work that consumes a chain's output WHILE the chain runs (another stream): returns once the chain's grid-barrier counter has reached
// `target` = steps * CTAs, i.e. every CTA has released its stores of that many steps (the acquire here + the kernel boundary after this kernel order the
// consumer's reads, TMA included, behind them).  One warp, one polling lane; a 20 s watchdog traps instead of hanging (the watcher may start
// well before the chain does).
__global__ void chain_progress_wait_kernel(const unsigned* __restrict__ gbar, unsigned target) {
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (unsigned spins = 1; ld_acquire_gpu(gbar) < target; ++spins) {
            __nanosleep(200);
            if ((spins & 255u) == 0 && clock64() - t0 > 40000000000LL) { printf("s2vt: chain progress watcher timed out (target %u, seen %u)\n", target, ld_acquire_gpu(gbar)); __trap(); }
        }
    }
}

// WS (weights stationary, a_rows <= 64 and K/KS <= 16 K-blocks): this CTA's weight slab (BN x K/KS, <= 64 KB) is fetched once
// and stays in shared memory for the whole chain; every A K-block of a step has its own stage (no ring, no `empty`
// barriers: the end-of-step barrier frees all stages), so all loads of a step are in flight at once.
// CX > 1 (weights stationary, KS == 1, a_rows % (8 CX) == 0): the CX CTAs of a cluster (consecutive column tiles) need the SAME activation
// rows; each fetches a_rows / CX of them per K-block and multicasts them to the whole cluster, so L2 serves every activation line
// ncta / CX times per step instead of ncta times (the step is bound by that broadcast, not by the per-SM fill rate: probe_chains.py).
// A peer may multicast into this CTA before it has armed its own `full` barrier for the step: the transaction count simply goes
// negative until the local expect_tx; it cannot run a phase ahead because peers pass the grid barrier only after this CTA's epilogue.
// (ChainAcc<BN>::N, gemm_tcgen05.cuh, is shared with the per-step kernels so both sum in the same order.)
// MMA2 (two accumulators): a SECOND issuing warp -- the last warp of the CTA -- takes the odd K-blocks (accumulator 1) while
// warp 1 takes the even ones (accumulator 0).  In situ one warp issues a tcgen05.mma every ~100 cycles (mbarrier wait, descriptors, four MMAs per K-block)
// against 62 in a tight loop (scripts/micro/mma_rate.cu); the K loop of a 64-row step is that issue chain.
template <int BN, class Epi, int KS, bool WS> struct ChainMma2 { static constexpr bool value = ChainAcc<BN>::N >= 2; };
template <int BN, class Epi, int KS, bool WS, int CX = 1>
__global__ void __launch_bounds__(Threads<BN, Epi>::N + (ChainMma2<BN, Epi, KS, WS>::value ? 32 * (ChainAcc<BN>::N - 1) : 0)) gemm_tc_chain_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                            int K, int a_rows, int a_row0, int a_row_stride,
                                                                            const typename Epi::Params* __restrict__ steps, int nsteps,
                                                                            unsigned* __restrict__ gbar, uint32_t fmt) {
    using C = Cfg<BN, Threads<BN, Epi>::N>;
    const uint32_t IDESC = C::IDESC & ~fmt;           // fmt: operand formats (FMT_A_F16 | FMT_B_F16 select fp16, gemm_tcgen05.cuh)
    static_assert(Epi::kDirect, "chain kernel: register epilogues only");
    static_assert(KS == 1 || KS == 4, "split-K cluster of 4 or none");
    static_assert(CX == 1 || (WS && KS == 1), "activation multicast: weights-stationary forward chains only");
    constexpr int NACC = ChainAcc<BN>::N;
    constexpr bool MMA2 = ChainMma2<BN, Epi, KS, WS>::value;
    constexpr int MMA2_WARP = Threads<BN, Epi>::N / 32;               // index of the second issuing warp (appended after the epilogue warps)
    constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation: power of two <= 512 columns");
    uint32_t crank = 0;
    if constexpr (CX > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int WS_KB = 16;                                          // K-blocks per CTA in weights-stationary mode
    constexpr int WS_A = 64 * 128;                                     // bytes of one A stage (64 rows x 64 bf16)
    constexpr int MAIN = WS ? WS_KB * (WS_A + C::B_BYTES) : C::PIPE_BYTES;
    constexpr int NBAR = WS ? WS_KB : C::STAGES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + MAIN);
    uint64_t* empty = full + NBAR;                                     // (ring mode only) / wfull in WS mode
    uint64_t* tmem_full = empty + NBAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* recv = reinterpret_cast<float*>(smem + MAIN + 512);          // [KS][32][BN] (KS > 1)
    unsigned char* wsm = smem + WS_KB * WS_A;                          // WS: weight slab [WS_KB][BN x 64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * C::BM, n0 = blockIdx.x * BN;
    const int rank = KS > 1 ? (int)blockIdx.z : 0;
    const int KBL = K / C::BK / KS, kb0 = rank * KBL;
    const unsigned ncta = gridDim.x * gridDim.y * gridDim.z;
    const uint32_t stage_tx = (uint32_t)(a_rows * 128 + C::B_BYTES);
    // debug probe (s2vt_debug_probe): CTA (0,0,0) records %globaltimer at the phase boundaries of every step, 8 slots per step
#ifdef S2VT_CHAIN_PROBE
    __shared__ unsigned long long* probe;
    if (threadIdx.x == 0) {
        probe = nullptr;
        if (g_probe && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            unsigned long long slot = atomicAdd(g_probe, (unsigned long long)nsteps);
            if (slot + nsteps < 4000) probe = g_probe + 8 * (slot + 1);
        }
    }
#endif

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < NBAR; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, MMA2 ? NACC : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if constexpr (WS) {   // the whole weight slab, once (weights never depend on the previous kernel)
            mbar_expect_tx(empty, (uint32_t)(KBL * C::B_BYTES));
            for (int i = 0; i < KBL; ++i) tma_load_2d_raw(wsm + i * C::B_BYTES, &mapB, empty, (kb0 + i) * C::BK, n0);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nacc = KBL < NACC ? KBL : NACC;      // accumulator tiles actually written by a step
    if constexpr (KS > 1 || CX > 1) cluster_sync_all();   // every CTA of the cluster has initialised its barriers
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int pre = KBL < C::STAGES ? KBL : C::STAGES;
    for (int s = 0; s < nsteps; ++s) {
        const int g0 = s * KBL;   // running K-block counter at the start of this step (pipeline phases continue across steps)
        // ---- weight tiles of the first stages: independent of the previous step, requested before the barrier
        if (!WS && warp == 0) {
            const bool leader = elect_one();
            for (int i = 0; i < pre; ++i) {
                const int g = g0 + i, st = g % C::STAGES;
                if (g >= C::STAGES) mbar_wait(empty + st, ((g / C::STAGES) - 1) & 1);
                if (leader) {
                    mbar_expect_tx(full + st, stage_tx);
                    tma_load_2d_raw(smem + st * C::STAGE_BYTES + C::A_BYTES, &mapB, full + st, (kb0 + i) * C::BK, n0);
                }
            }
            __syncwarp();
        }
        // ---- dependency on the previous step (s == 0: on the previous kernel)
        if (s == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
        } else {
            __syncthreads();                       // this CTA's stores of step s-1 are issued
            if (threadIdx.x == 0) {
                CHAIN_PROBE(if (probe) probe[8 * s + 0] = gtimer());
                // release: orders this CTA's stores (made visible to thread 0 by the bar.sync above, cumulativity) before the arrival
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(gbar), "r"(1u) : "memory");
                const unsigned target = (unsigned)s * ncta;
                long long t0 = clock64();
                for (unsigned spins = 1; ld_acquire_gpu(gbar) < target; ++spins) {
                    if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) { printf("s2vt: grid barrier timed out (step %d, block %d,%d,%d)\n", s, blockIdx.x, blockIdx.y, blockIdx.z); __trap(); }
                }
                CHAIN_PROBE(if (probe) { probe[8 * s + 1] = gtimer(); probe[8 * s + 6] = ((unsigned long long)(BN + 1000 * KS + (WS ? 100000 : 0)) << 32) | (unsigned)K; });
            }
            // (Letting the non-producer warps run ahead of the grid barrier -- bar.arrive instead of the bar.sync below -- was measured: 1.5 %
            // slower; their early operand prefetch competes with the activation fetch that the whole grid is waiting for.)
            __syncthreads();
        }
        const typename Epi::Params& ep = steps[s];
        const int arow = a_row0 + s * a_row_stride + m0;

        if (warp == 0) {
            const bool leader = elect_one();
            if (leader) asm volatile("fence.proxy.async;" ::: "memory");   // rows written through the generic proxy by other SMs are read by TMA
            if constexpr (WS && CX > 1) {
                const int rows_per = a_rows / CX;          // box height of mapA in this mode
                if (leader)
                    for (int i = 0; i < KBL; ++i) {
                        mbar_expect_tx(full + i, (uint32_t)(a_rows * 128));
                        tma_load_2d_mc(smem + i * WS_A + crank * rows_per * 128, &mapA, full + i, (kb0 + i) * C::BK, arow + (int)crank * rows_per, (uint16_t)((1u << CX) - 1u));
                    }
            } else if constexpr (WS) {
                if (leader)      // (3-D boxes fetching 4 K-block tiles per instruction were measured: no faster, the step is MMA-issue bound)
                    for (int i = 0; i < KBL; ++i) {
                        mbar_expect_tx(full + i, (uint32_t)(a_rows * 128));
                        tma_load_2d_raw(smem + i * WS_A, &mapA, full + i, (kb0 + i) * C::BK, arow);
                    }
            } else {
                for (int i = 0; i < KBL; ++i) {
                    const int g = g0 + i, st = g % C::STAGES;
                    unsigned char* a = smem + st * C::STAGE_BYTES;
                    if (i >= pre) {
                        mbar_wait(empty + st, ((g / C::STAGES) - 1) & 1);
                        if (leader) {
                            mbar_expect_tx(full + st, stage_tx);
                            tma_load_2d_raw(a + C::A_BYTES, &mapB, full + st, (kb0 + i) * C::BK, n0);
                        }
                    }
                    if (leader) tma_load_2d_raw(a, &mapA, full + st, (kb0 + i) * C::BK, arow);
                }
            }
            __syncwarp();
        } else if (warp == 1 || (MMA2 && warp >= MMA2_WARP)) {
            const bool leader = elect_one();
            if constexpr (WS) {
                if (s == 0) mbar_wait(empty, 0);           // weight slab has landed
                const int i0 = (MMA2 && warp != 1) ? warp - MMA2_WARP + 1 : 0, istep = MMA2 ? NACC : 1;      // MMA2: this warp's K-blocks (i mod NACC)
                for (int i = i0; i < KBL; i += istep) {
                    mbar_wait(full + i, s & 1);
                    if (i == i0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");   // once per step: orders the MMAs after the previous epilogue's TMEM reads
                    CHAIN_PROBE(if (probe && i == 0 && leader) probe[8 * s + 2] = gtimer()); CHAIN_PROBE(if (probe && i == KBL - 1 && leader) probe[8 * s + 7] = gtimer());
                    const uint64_t adesc = make_desc(smem_u32(smem + i * WS_A)), bdesc = make_desc(smem_u32(wsm + i * C::B_BYTES));
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < C::BK / 16; ++k) mma_bf16(tmem_base + (uint32_t)((i % NACC) * BN), adesc + 2 * k, bdesc + 2 * k, IDESC, i >= NACC || k != 0);
                    }
                }
            } else {
                const int i0 = (MMA2 && warp != 1) ? warp - MMA2_WARP + 1 : 0, istep = MMA2 ? NACC : 1;
                for (int i = i0; i < KBL; i += istep) {
                    const int g = g0 + i, st = g % C::STAGES;
                    mbar_wait(full + st, (g / C::STAGES) & 1);
                    if (i == i0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    CHAIN_PROBE(if (probe && i == 0 && leader) probe[8 * s + 2] = gtimer()); CHAIN_PROBE(if (probe && i == KBL - 1 && leader) probe[8 * s + 7] = gtimer());
                    const uint32_t a = smem_u32(smem + st * C::STAGE_BYTES);
                    const uint64_t adesc = make_desc(a), bdesc = make_desc(a + C::A_BYTES);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < C::BK / 16; ++k) mma_bf16(tmem_base + (uint32_t)((i % NACC) * BN), adesc + 2 * k, bdesc + 2 * k, IDESC, i >= NACC || k != 0);
                        mma_commit(empty + st);
                    }
                }
            }
            if (leader) {
                mma_commit(tmem_full);
                CHAIN_PROBE(if (probe) probe[8 * s + 3] = gtimer());
            }
            __syncwarp();
        } else {
            const int e = warp - 2, q = warp & 3;
            const int row = q * 32 + lane;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            if constexpr (KS == 1) {
                const int c0 = (e >> 2) * 32;
                typename Epi::Pre prf;
                Epi::prefetch(ep, m0 + row, n0 + c0, prf);
                mbar_wait(tmem_full, s & 1);
                CHAIN_PROBE(if (probe && threadIdx.x == 128) probe[8 * s + 4] = gtimer());
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                for (int a = 1; a < nacc; ++a) {
                    float w[32];
                    tmem_ld32(trow + (uint32_t)(a * BN + c0), w);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += w[j];
                }
                direct_chunk<Epi>(ep, m0 + row, n0 + c0, v, prf, false);
                CHAIN_PROBE(if (probe && threadIdx.x == 128) probe[8 * s + 5] = gtimer());
            } else {
                constexpr int UPR = BN / 8;
                const int t = threadIdx.x - 64;
                const int frow = t / UPR, fc8 = t % UPR;
                typename Epi::Pre prf;
                Epi::prefetch(ep, m0 + 32 * rank + frow, n0 + 8 * fc8, prf, 1);
                mbar_wait(tmem_full, s & 1);
                CHAIN_PROBE(if (probe && threadIdx.x == 128) probe[8 * s + 4] = gtimer());
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
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
                    // of the peer's shared memory (row-major tiles made every lane hit its own 512-byte-strided row)
                    const uint32_t base = cluster_map(smem_u32(recv), (uint32_t)q) + (uint32_t)(((rank * (BN / 4) + (c0 >> 2)) * 32) * 16);
#pragma unroll
                    for (int j = 0; j < 8; ++j)      // row slot XOR-ed with the 8-unit group index: the reducer's 16 lanes of a row then spread over the banks
                        st_cluster_f4(base + j * 512 + ((lane ^ ((((c0 >> 2) + j) >> 1) & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                Epi::prefetch(ep, m0 + 32 * rank + frow, n0 + 8 * fc8, prf, 2);
#ifdef S2VT_CHAIN_PROBE_EPI
                CHAIN_PROBE(if (probe && threadIdx.x == 128) probe[8 * s + 7] = gtimer());
#endif
                cluster_sync_all();
#ifdef S2VT_CHAIN_PROBE_EPI
                CHAIN_PROBE(if (probe && threadIdx.x == 128) probe[8 * s + 2] = gtimer());
#endif
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int src = 0; src < KS; ++src) {
                    const float4* rp = reinterpret_cast<const float4*>(recv) + (src * (BN / 4) + 2 * fc8) * 32 + (frow ^ (fc8 & 7));
                    const float4 x0 = rp[0], x1 = rp[32];
                    acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w; acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
                }
                Epi::direct(ep, m0 + 32 * rank + frow, n0 + 8 * fc8, acc, prf);
                CHAIN_PROBE(if (probe && threadIdx.x == 128) probe[8 * s + 5] = gtimer());
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        if constexpr (KS > 1) {
            if (warp < 2 || (MMA2 && warp >= MMA2_WARP)) cluster_sync_all();      // pairs with the exchange barrier of the epilogue warps
            // No second cluster barrier: a peer writes into this CTA's recv buffers again only after the NEXT grid barrier, which it
            // passes only after every CTA -- this one included -- has finished the epilogue that reads them.
        }
    }
    __syncthreads();
    if constexpr (KS > 1) cluster_sync_all();      // nobody leaves while a peer may still be reading its exchange buffers
    if constexpr (CX > 1) cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Host launcher.  A: one buffer of `a_total_rows` rows (row stride lda); step s reads rows [a_row0 + s * a_row_stride, + M).
// Returns cudaErrorLaunchOutOfResources (without launching) if the grid cannot be co-resident.
template <int BN, class Epi, int KS, bool WS = false, int CX = 1>
inline cudaError_t launch_chain(MapCache& cache, cudaStream_t st, const bf16* A, int lda, int a_total_rows, int a_row0, int a_row_stride, const bf16* B,
                                int ldb, int M, int N, int K, const typename Epi::Params* steps_dev, int nsteps, unsigned* gbar, bool pdl, uint32_t fmt = 0) {
    constexpr int NT = Threads<BN, Epi>::N;
    constexpr int NTL = NT + (ChainMma2<BN, Epi, KS, WS>::value ? 32 * (ChainAcc<BN>::N - 1) : 0);       // + the additional issuing warps
    using C = Cfg<BN, NT>;
    constexpr int SMEM = (WS ? 16 * (64 * 128 + C::B_BYTES) + 512 + 1024 : C::SMEM_BYTES + 512) + (KS > 1 ? KS * 32 * BN * 4 : 0);
    if (WS && (M > 64 || K / BK / KS > 16)) return cudaErrorLaunchOutOfResources;
    if (cache.size() > 32768) cache.clear();
    const int a_rows = M <= 64 ? ((M + 7) & ~7) : BM;
    if (CX > 1 && (a_rows % (8 * CX) != 0 || (N / BN) % CX != 0)) return cudaErrorLaunchOutOfResources;
    const CUtensorMap* ma = get_map(cache, A, a_total_rows, K, lda, a_rows / CX);
    const CUtensorMap* mb = get_map(cache, B, N, K, ldb, BN);
    if (!ma || !mb) return cudaErrorInvalidValue;
    auto kern = gemm_tc_chain_kernel<BN, Epi, KS, WS, CX>;
    static int max_ctas = -1;
    if (max_ctas < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NTL, SMEM);
        if (e != cudaSuccess) return e;
        max_ctas = per_sm * sms;
        if (CX > 1) {   // clusters must fit inside GPCs: ask how many can be resident at once
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(N / BN, (M + BM - 1) / BM, KS); q.blockDim = dim3(NTL); q.dynamicSmemBytes = SMEM;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = CX; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = KS;
            q.attrs = qa; q.numAttrs = 1;
            int ncl = 0;
            e = cudaOccupancyMaxActiveClusters(&ncl, kern, &q);
            if (e != cudaSuccess) { (void)cudaGetLastError(); ncl = 0; }
            max_ctas = ncl * CX * KS;
        }
    }
    dim3 grid(N / BN, (M + BM - 1) / BM, KS);
    static const bool debug = getenv("S2VT_DEBUG_CHAIN") != nullptr;
    if (debug) fprintf(stderr, "s2vt chain: BN=%d KS=%d WS=%d CX=%d rows=%d N=%d K=%d steps=%d grid=%u co-resident limit=%d\n", BN, KS, (int)WS, CX, M, N, K, nsteps,
                       grid.x * grid.y * grid.z, max_ctas);
    if ((int)(grid.x * grid.y * grid.z) > max_ctas) return cudaErrorLaunchOutOfResources;
    cudaError_t e = cudaMemsetAsync(gbar, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NTL);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (KS > 1 || CX > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CX; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = KS;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, *ma, *mb, K, a_rows, a_row0, a_row_stride, steps_dev, nsteps, gbar, fmt);
}

}  // namespace tc
