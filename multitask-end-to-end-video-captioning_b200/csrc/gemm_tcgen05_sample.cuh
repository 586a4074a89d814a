// Persistent sampling chain: ONE launch walks every decode step of a rollout (build_multinomial_sampler :294-339 / build_sampler
// :342-391), two phases per step separated by grid barriers
//     cell phase   h2', c2' = BasicLSTMCell(h2 . W2h + G2x[t] + Etab[word])        128 x 128 tiles, register epilogue (EpiLstmFwd)
//     pick phase   word     = arg-max over (h2' . Wo + bo [+ Gumbel noise])        128 x 256 tiles, staged epilogue  (EpiLogitsPick)
// instead of two kernel launches per step.  The phases are the per-step kernels of gemm_tcgen05.cuh (same TMA / tcgen05 / TMEM
// pipeline, same epilogue functors, same accumulation order -> identical words); what disappears is the pair of kernel boundaries
// per step (grid drain, completion flush, dependency release, prologue): a grid barrier costs ~1.5 us, a boundary 3-4 us plus the
// ramp of the next kernel.
//
// One CTA per SM (co-residency checked by the host).  CTA c owns cell tile c (if c < #cell tiles) and pick tile c (if c < #pick
// tiles); both pipelines live in the same shared memory (they never overlap in time: a phase ends with every MMA retired and a
// __syncthreads) with separate mbarrier rings whose phase counters run across the steps.  Weight tiles of the next phase are requested
// before its grid barrier (they never depend on it).  Memory ordering across a barrier is that of gemm_tcgen05_chain.cuh.
#pragma once
#include "gemm_tcgen05_chain.cuh"

namespace tc {

template <class CellEpi, class PickEpi>
__global__ void __launch_bounds__(576) sample_chain_kernel(const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapH1,
                                                           const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWo,
                                                           int K, int n_mt, int n_cell_n, int n_pick_n,
                                                           const typename CellEpi::Params* __restrict__ cell_steps,
                                                           const typename PickEpi::Params* __restrict__ pick_steps, int nsteps,
                                                           unsigned* __restrict__ gbar, uint32_t fmt) {
    static_assert(CellEpi::kDirect && !PickEpi::kDirect, "cell: register epilogue, pick: staged epilogue");
    constexpr int NT = 576;
    using CC = Cfg<128, NT>;           // cell tiles
    using CP = Cfg<256, NT>;           // pick tiles
    static_assert(Threads<128, CellEpi>::N == NT && Threads<256, PickEpi>::N == NT, "both phases run on 2 role warps + 16 epilogue warps");
    constexpr int MAIN = CP::PIPE_BYTES > CC::PIPE_BYTES ? (CP::PIPE_BYTES > CP::EPI_BYTES ? CP::PIPE_BYTES : CP::EPI_BYTES)
                                                         : (CC::PIPE_BYTES > CP::EPI_BYTES ? CC::PIPE_BYTES : CP::EPI_BYTES);
    constexpr int TMEM_COLS = 256;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* fullC = reinterpret_cast<uint64_t*>(smem + MAIN);
    uint64_t* emptyC = fullC + CC::STAGES;
    uint64_t* fullP = emptyC + CC::STAGES;
    uint64_t* emptyP = fullP + CP::STAGES;
    uint64_t* tmem_full = emptyP + CP::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* Cs = reinterpret_cast<float*>(smem);      // pick tile, aliases the pipelines once every MMA of the phase has retired

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const unsigned ncta = gridDim.x;
    const bool has_cell = cta < n_mt * n_cell_n, has_pick = cta < n_mt * n_pick_n;
    const int m0c = (cta / n_cell_n) * BM, n0c = (cta % n_cell_n) * 128;
    const int m0p = (cta / n_pick_n) * BM, n0p = (cta % n_pick_n) * 256;
    const int KBL = K / BK;
    constexpr uint32_t txC = (uint32_t)CC::STAGE_BYTES, txP = (uint32_t)CP::STAGE_BYTES;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWo) : "memory");
        for (int s = 0; s < CC::STAGES; ++s) { mbar_init(fullC + s, 1); mbar_init(emptyC + s, 1); }
        for (int s = 0; s < CP::STAGES; ++s) { mbar_init(fullP + s, 1); mbar_init(emptyP + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int preC = KBL < CC::STAGES ? KBL : CC::STAGES, preP = KBL < CP::STAGES ? KBL : CP::STAGES;
    unsigned nbar = 0;        // grid barriers passed so far
    int uses = 0;             // completed uses of tmem_full by this CTA (its wait parity)

    // grid barrier: every thread calls it; thread 0 arrives (release) and spins (acquire), the bar.syncs extend the ordering to the CTA
    auto grid_barrier = [&]() {
        __syncthreads();
        ++nbar;
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(gbar), "r"(1u) : "memory");
            const unsigned target = nbar * ncta;
            long long t0 = clock64();
            for (unsigned spins = 1; ld_acquire_gpu(gbar) < target; ++spins) {
                if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) { printf("s2vt: sampling-chain grid barrier %u timed out (block %d)\n", nbar, blockIdx.x); __trap(); }
            }
        }
        __syncthreads();
    };

    for (int s = 0; s < nsteps; ++s) {
        // ================================================ cell phase ================================================
        const int gC0 = s * KBL;
        if (has_cell && warp == 0) {       // W2h tiles of the first stages, before the barrier
            const bool leader = elect_one();
            for (int i = 0; i < preC; ++i) {
                const int g = gC0 + i, st = g % CC::STAGES;
                if (g >= CC::STAGES) mbar_wait(emptyC + st, ((g / CC::STAGES) - 1) & 1);
                if (leader) {
                    mbar_expect_tx(fullC + st, txC);
                    tma_load_2d_raw(smem + st * CC::STAGE_BYTES + CC::A_BYTES, &mapWh, fullC + st, i * BK, n0c);
                }
            }
            __syncwarp();
        }
        if (s == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        else grid_barrier();               // the words of step s-1 are picked
        if (has_cell) {
            const typename CellEpi::Params& ep = cell_steps[s];
            const CUtensorMap* mapA = (s & 1) ? &mapH1 : &mapH0;
            if (warp == 0) {
                const bool leader = elect_one();
                if (leader) asm volatile("fence.proxy.async;" ::: "memory");
                for (int i = 0; i < KBL; ++i) {
                    const int g = gC0 + i, st = g % CC::STAGES;
                    unsigned char* a = smem + st * CC::STAGE_BYTES;
                    if (i >= preC) {
                        mbar_wait(emptyC + st, ((g / CC::STAGES) - 1) & 1);
                        if (leader) {
                            mbar_expect_tx(fullC + st, txC);
                            tma_load_2d_raw(a + CC::A_BYTES, &mapWh, fullC + st, i * BK, n0c);
                        }
                    }
                    if (leader) tma_load_2d_raw(a, mapA, fullC + st, i * BK, m0c);
                }
                __syncwarp();
            } else if (warp == 1) {
                const bool leader = elect_one();
                for (int i = 0; i < KBL; ++i) {
                    const int g = gC0 + i, st = g % CC::STAGES;
                    mbar_wait(fullC + st, (g / CC::STAGES) & 1);
                    if (i == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a = smem_u32(smem + st * CC::STAGE_BYTES);
                    const uint64_t adesc = make_desc(a), bdesc = make_desc(a + CC::A_BYTES);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, (CC::IDESC & ~fmt), i > 0 || k != 0);
                        mma_commit(emptyC + st);
                    }
                }
                if (leader) mma_commit(tmem_full);
                __syncwarp();
            } else {
                const int e = warp - 2, q = warp & 3;
                const int row = q * 32 + lane;
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
                const int c0 = (e >> 2) * 32;
                typename CellEpi::Pre prf;
                CellEpi::prefetch(ep, m0c + row, n0c + c0, prf);
                mbar_wait(tmem_full, uses & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float v[32];
                tmem_ld32(trow + (uint32_t)c0, v);
                direct_chunk<CellEpi>(ep, m0c + row, n0c + c0, v, prf, false);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            ++uses;
        }
        // ================================================ pick phase ================================================
        __syncthreads();                   // the cell pipeline is drained: its shared memory may take the Wo tiles
        const int gP0 = s * KBL;
        if (has_pick && warp == 0) {
            const bool leader = elect_one();
            for (int i = 0; i < preP; ++i) {
                const int g = gP0 + i, st = g % CP::STAGES;
                if (g >= CP::STAGES) mbar_wait(emptyP + st, ((g / CP::STAGES) - 1) & 1);
                if (leader) {
                    mbar_expect_tx(fullP + st, txP);
                    tma_load_2d_raw(smem + st * CP::STAGE_BYTES + CP::A_BYTES, &mapWo, fullP + st, i * BK, n0p);
                }
            }
            __syncwarp();
        }
        grid_barrier();                    // h2' of every row is written
        if (has_pick) {
            const typename PickEpi::Params& ep = pick_steps[s];
            const CUtensorMap* mapA = (s & 1) ? &mapH0 : &mapH1;     // the buffer the cell phase just wrote
            if (warp == 0) {
                const bool leader = elect_one();
                if (leader) asm volatile("fence.proxy.async;" ::: "memory");
                for (int i = 0; i < KBL; ++i) {
                    const int g = gP0 + i, st = g % CP::STAGES;
                    unsigned char* a = smem + st * CP::STAGE_BYTES;
                    if (i >= preP) {
                        mbar_wait(emptyP + st, ((g / CP::STAGES) - 1) & 1);
                        if (leader) {
                            mbar_expect_tx(fullP + st, txP);
                            tma_load_2d_raw(a + CP::A_BYTES, &mapWo, fullP + st, i * BK, n0p);
                        }
                    }
                    if (leader) tma_load_2d_raw(a, mapA, fullP + st, i * BK, m0p);
                }
                __syncwarp();
            } else if (warp == 1) {
                const bool leader = elect_one();
                for (int i = 0; i < KBL; ++i) {
                    const int g = gP0 + i, st = g % CP::STAGES;
                    mbar_wait(fullP + st, (g / CP::STAGES) & 1);
                    if (i == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a = smem_u32(smem + st * CP::STAGE_BYTES);
                    const uint64_t adesc = make_desc(a), bdesc = make_desc(a + CP::A_BYTES);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, (CP::IDESC & ~fmt), i > 0 || k != 0);
                        mma_commit(emptyP + st);
                    }
                }
                if (leader) mma_commit(tmem_full);
                __syncwarp();
            } else {
                const int e = warp - 2, q = warp & 3;
                const int row = q * 32 + lane;
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
                mbar_wait(tmem_full, uses & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int cbeg = (e >> 2) * 64;      // 16 epilogue warps: 4 lane quarters x 4 column ranges of 64
#pragma unroll
                for (int c0 = cbeg; c0 < cbeg + 64; c0 += 32) {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(Cs + row * CP::LDC + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            ++uses;
            __syncthreads();
            PickEpi::template apply<CP>(ep, Cs, m0p, n0p);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of Cs before the TMA writes that reuse it
        }
        // the next cell phase starts with __syncthreads inside grid_barrier(): Cs is no longer read when its weight prefetch lands
        __syncthreads();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Overlapped variant: h2(s+1) . W2h does not depend on the word chosen in step s (only the cell epilogue does, through Etab[word]), so the
// TMA / MMA warps run the cell GEMM of step s+1 -- two pipeline stages beside the staged pick tile, accumulator in its own TMEM columns
// [256, 384) -- while the 16 epilogue warps draw the Gumbel noise of step s.  Per step: barrier A (words visible) -> cell epilogue ->
// barrier B (h2' visible; Wo tiles requested under it) -> pick MMAs -> { pick epilogue || cell MMAs of the next step }.
template <class CellEpi, class PickEpi>
__global__ void __launch_bounds__(576) sample_chain_ovl_kernel(const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapH1,
                                                               const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWo,
                                                               int K, int n_mt, int n_cell_n, int n_pick_n,
                                                               const typename CellEpi::Params* __restrict__ cell_steps,
                                                               const typename PickEpi::Params* __restrict__ pick_steps, int nsteps,
                                                               unsigned* __restrict__ gbar, uint32_t fmt) {
    static_assert(CellEpi::kDirect && !PickEpi::kDirect, "cell: register epilogue, pick: staged epilogue");
    constexpr int NT = 576;
    using CC = Cfg<128, NT>;
    using CP = Cfg<256, NT>;
    constexpr int SST = 2;                               // stages of the cell ring beside the pick tile
    constexpr int SIDE = CP::EPI_BYTES;                  // byte offset of that ring
    static_assert(SIDE % 1024 == 0, "swizzled stages need 1024-byte alignment");
    constexpr int MAIN = SIDE + SST * CC::STAGE_BYTES > CP::PIPE_BYTES ? SIDE + SST * CC::STAGE_BYTES : CP::PIPE_BYTES;
    constexpr int TMEM_COLS = 512;
    constexpr uint32_t CELL_COL = 256;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* fullS = reinterpret_cast<uint64_t*>(smem + MAIN);
    uint64_t* emptyS = fullS + SST;
    uint64_t* fullP = emptyS + SST;
    uint64_t* emptyP = fullP + CP::STAGES;
    uint64_t* tfullC = emptyP + CP::STAGES;
    uint64_t* tfullP = tfullC + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfullP + 1);
    float* Cs = reinterpret_cast<float*>(smem);
    unsigned char* side = smem + SIDE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const unsigned ncta = gridDim.x;
    const bool has_cell = cta < n_mt * n_cell_n, has_pick = cta < n_mt * n_pick_n;
    const int m0c = (cta / n_cell_n) * BM, n0c = (cta % n_cell_n) * 128;
    const int m0p = (cta / n_pick_n) * BM, n0p = (cta % n_pick_n) * 256;
    const int KBL = K / BK;
    constexpr uint32_t txC = (uint32_t)CC::STAGE_BYTES, txP = (uint32_t)CP::STAGE_BYTES;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWo) : "memory");
        for (int s = 0; s < SST; ++s) { mbar_init(fullS + s, 1); mbar_init(emptyS + s, 1); }
        for (int s = 0; s < CP::STAGES; ++s) { mbar_init(fullP + s, 1); mbar_init(emptyP + s, 1); }
        mbar_init(tfullC, 1); mbar_init(tfullP, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int preP = KBL < CP::STAGES ? KBL : CP::STAGES;
    unsigned nbar = 0;
    // debug probe (scripts/probe_sample_chain.py, -DS2VT_CHAIN_PROBE only): CTA 0 records %globaltimer at the phase boundaries of every step
#ifdef S2VT_CHAIN_PROBE
    __shared__ unsigned long long* probe;
    if (threadIdx.x == 0) {
        probe = nullptr;
        if (g_probe && blockIdx.x == 0) {
            unsigned long long slot = atomicAdd(g_probe, (unsigned long long)nsteps);
            if (slot + nsteps < 4000) probe = g_probe + 8 * (slot + 1);
        }
    }
    __syncthreads();
#endif

    // thread 0: arrive at grid barrier number nbar (release) and wait for everybody (acquire); callers bracket it with __syncthreads
    auto arrive_and_spin = [&]() {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(gbar), "r"(1u) : "memory");
        const unsigned target = nbar * ncta;
        long long t0 = clock64();
        for (unsigned spins = 1; ld_acquire_gpu(gbar) < target; ++spins) {
            if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) { printf("s2vt: sampling-chain grid barrier %u timed out (block %d)\n", nbar, blockIdx.x); __trap(); }
        }
    };
    // warps 0 / 1: TMA loads and MMAs of  h2(s) . W2h  ->  TMEM columns [CELL_COL, CELL_COL + 128)
    auto cell_mma = [&](int s) {
        const int g0 = s * KBL;
        const CUtensorMap* mapA = (s & 1) ? &mapH1 : &mapH0;
        if (warp == 0) {
            const bool leader = elect_one();
            if (leader) asm volatile("fence.proxy.async;" ::: "memory");
            for (int i = 0; i < KBL; ++i) {
                const int g = g0 + i, st = g % SST;
                if (g >= SST) mbar_wait(emptyS + st, ((g / SST) - 1) & 1);
                if (leader) {
                    unsigned char* a = side + st * CC::STAGE_BYTES;
                    mbar_expect_tx(fullS + st, txC);
                    tma_load_2d_raw(a + CC::A_BYTES, &mapWh, fullS + st, i * BK, n0c);
                    tma_load_2d_raw(a, mapA, fullS + st, i * BK, m0c);
                }
            }
            __syncwarp();
        } else {
            const bool leader = elect_one();
            for (int i = 0; i < KBL; ++i) {
                const int g = g0 + i, st = g % SST;
                mbar_wait(fullS + st, (g / SST) & 1);
                if (i == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a = smem_u32(side + st * CC::STAGE_BYTES);
                const uint64_t adesc = make_desc(a), bdesc = make_desc(a + CC::A_BYTES);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base + CELL_COL, adesc + 2 * k, bdesc + 2 * k, (CC::IDESC & ~fmt), i > 0 || k != 0);
                    mma_commit(emptyS + st);
                }
            }
            if (leader) mma_commit(tfullC);
            __syncwarp();
        }
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (has_cell && warp < 2) cell_mma(0);
    for (int s = 0; s < nsteps; ++s) {
        // ---- barrier A: the words of step s-1 are picked
        if (s > 0) {
            __syncthreads();
            ++nbar;
            CHAIN_PROBE(if (probe && threadIdx.x == 0) probe[8 * s + 0] = gtimer());
            if (threadIdx.x == 0) arrive_and_spin();
            __syncthreads();
        }
        CHAIN_PROBE(if (probe && threadIdx.x == 0) probe[8 * s + 1] = gtimer());
        // ---- cell epilogue of step s (its MMAs ran under the previous pick epilogue)
        if (has_cell && warp >= 2) {
            const typename CellEpi::Params& ep = cell_steps[s];
            const int e = warp - 2, q = warp & 3;
            const int row = q * 32 + lane;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + CELL_COL;
            const int c0 = (e >> 2) * 32;
            typename CellEpi::Pre prf;
            CellEpi::prefetch(ep, m0c + row, n0c + c0, prf);
            mbar_wait(tfullC, s & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[32];
            tmem_ld32(trow + (uint32_t)c0, v);
            direct_chunk<CellEpi>(ep, m0c + row, n0c + c0, v, prf, false);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        // ---- barrier B: h2' of every row is written; the first Wo tiles are requested under it
        __syncthreads();
        CHAIN_PROBE(if (probe && threadIdx.x == 0) probe[8 * s + 2] = gtimer());
        const int gP0 = s * KBL;
        if (has_pick && warp == 0) {
            const bool leader = elect_one();
            for (int i = 0; i < preP; ++i) {
                const int g = gP0 + i, st = g % CP::STAGES;
                if (g >= CP::STAGES) mbar_wait(emptyP + st, ((g / CP::STAGES) - 1) & 1);
                if (leader) {
                    mbar_expect_tx(fullP + st, txP);
                    tma_load_2d_raw(smem + st * CP::STAGE_BYTES + CP::A_BYTES, &mapWo, fullP + st, i * BK, n0p);
                }
            }
            __syncwarp();
        }
        ++nbar;
        if (threadIdx.x == 0) arrive_and_spin();
        __syncthreads();
        CHAIN_PROBE(if (probe && threadIdx.x == 0) probe[8 * s + 3] = gtimer());
        // ---- pick MMAs of step s
        if (has_pick) {
            const CUtensorMap* mapA = (s & 1) ? &mapH0 : &mapH1;     // the buffer the cell epilogue just wrote
            if (warp == 0) {
                const bool leader = elect_one();
                if (leader) asm volatile("fence.proxy.async;" ::: "memory");
                for (int i = 0; i < KBL; ++i) {
                    const int g = gP0 + i, st = g % CP::STAGES;
                    unsigned char* a = smem + st * CP::STAGE_BYTES;
                    if (i >= preP) {
                        mbar_wait(emptyP + st, ((g / CP::STAGES) - 1) & 1);
                        if (leader) {
                            mbar_expect_tx(fullP + st, txP);
                            tma_load_2d_raw(a + CP::A_BYTES, &mapWo, fullP + st, i * BK, n0p);
                        }
                    }
                    if (leader) tma_load_2d_raw(a, mapA, fullP + st, i * BK, m0p);
                }
                __syncwarp();
            } else if (warp == 1) {
                const bool leader = elect_one();
                for (int i = 0; i < KBL; ++i) {
                    const int g = gP0 + i, st = g % CP::STAGES;
                    mbar_wait(fullP + st, (g / CP::STAGES) & 1);
                    if (i == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a = smem_u32(smem + st * CP::STAGE_BYTES);
                    const uint64_t adesc = make_desc(a), bdesc = make_desc(a + CP::A_BYTES);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, (CP::IDESC & ~fmt), i > 0 || k != 0);
                        mma_commit(emptyP + st);
                    }
                }
                if (leader) mma_commit(tfullP);
                __syncwarp();
            } else {
                const int e = warp - 2, q = warp & 3;
                const int row = q * 32 + lane;
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
                mbar_wait(tfullP, s & 1);
                CHAIN_PROBE(if (probe && threadIdx.x == 64) probe[8 * s + 4] = gtimer());
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int cbeg = (e >> 2) * 64;
#pragma unroll
                for (int c0 = cbeg; c0 < cbeg + 64; c0 += 32) {
                    float v[32];
                    tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(Cs + row * CP::LDC + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        }
        __syncthreads();                   // pick tile staged, pick pipeline drained
        CHAIN_PROBE(if (probe && threadIdx.x == 0) probe[8 * s + 5] = gtimer());
        // ---- pick epilogue of step s (epilogue warps)  ||  cell MMAs of step s+1 (role warps)
        if (warp < 2) {
            if (has_cell && s + 1 < nsteps) cell_mma(s + 1);
            CHAIN_PROBE(if (probe && threadIdx.x == 32) probe[8 * s + 7] = gtimer());      // MMAs issued (their completion is tfullC)
        } else if (has_pick) {
            PickEpi::template apply_on<CP, 64, NT - 64>(pick_steps[s], Cs, m0p, n0p);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of Cs before the TMA writes that reuse it
            CHAIN_PROBE(if (probe && threadIdx.x == 64) probe[8 * s + 6] = gtimer());
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Host launcher.  H0 / H1: the ping-pong hidden-state buffers [rows, K] (step s reads H[s & 1], writes H[(s + 1) & 1] through the cell
// epilogue's h_out); Wh [Ncell, K], Wo [Npick, K] K-major.  cudaErrorLaunchOutOfResources (nothing launched) if the shape does not fit.
template <class CellEpi, class PickEpi, bool OVL = true>
inline cudaError_t launch_sample_chain(MapCache& cache, cudaStream_t st, const bf16* H0, const bf16* H1, int ldh, int rows, const bf16* Wh, int ldwh, int Ncell,
                                       const bf16* Wo, int ldwo, int Npick, int K, const typename CellEpi::Params* cell_dev,
                                       const typename PickEpi::Params* pick_dev, int nsteps, unsigned* gbar, bool pdl, uint32_t fmt = 0) {
    constexpr int NT = 576;
    using CC = Cfg<128, NT>;
    using CP = Cfg<256, NT>;
    constexpr int MAIN = CP::PIPE_BYTES > CC::PIPE_BYTES ? (CP::PIPE_BYTES > CP::EPI_BYTES ? CP::PIPE_BYTES : CP::EPI_BYTES)
                                                         : (CC::PIPE_BYTES > CP::EPI_BYTES ? CC::PIPE_BYTES : CP::EPI_BYTES);
    constexpr int MAIN_OVL = CP::EPI_BYTES + 2 * CC::STAGE_BYTES > CP::PIPE_BYTES ? CP::EPI_BYTES + 2 * CC::STAGE_BYTES : CP::PIPE_BYTES;
    constexpr int SMEM = (OVL ? MAIN_OVL : MAIN) + 512 + 1024;
    if (rows <= 0 || Ncell % 128 != 0 || Npick % 256 != 0 || K % BK != 0 || nsteps <= 0) return cudaErrorLaunchOutOfResources;
    const int n_mt = (rows + BM - 1) / BM, n_cell_n = Ncell / 128, n_pick_n = Npick / 256;
    const int grid = n_mt * (n_cell_n > n_pick_n ? n_cell_n : n_pick_n);
    if (cache.size() > 32768) cache.clear();
    const CUtensorMap* mh0 = get_map(cache, H0, rows, K, ldh, BM);
    const CUtensorMap* mh1 = get_map(cache, H1, rows, K, ldh, BM);
    const CUtensorMap* mwh = get_map(cache, Wh, Ncell, K, ldwh, 128);
    const CUtensorMap* mwo = get_map(cache, Wo, Npick, K, ldwo, 256);
    if (!mh0 || !mh1 || !mwh || !mwo) return cudaErrorInvalidValue;
    auto kern = OVL ? sample_chain_ovl_kernel<CellEpi, PickEpi> : sample_chain_kernel<CellEpi, PickEpi>;
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
    if (grid > max_ctas) return cudaErrorLaunchOutOfResources;
    cudaError_t e = cudaMemsetAsync(gbar, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
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
    return cudaLaunchKernelEx(&cfg, kern, *mh0, *mh1, *mwh, *mwo, K, n_mt, n_cell_n, n_pick_n, cell_dev, pick_dev, nsteps, gbar, fmt);
}

}  // namespace tc
