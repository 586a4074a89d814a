// Data-parallel gradient exchange over NVLink peer memory (one process per GPU, all GPUs behind one NVSwitch).
//
// The path's only exchange step is the sum of the flat fp32 gradient block (+ aux slots) over the ranks once per iteration
// (trainer.allreduce_gradients).  NCCL does it in 0.39 ms for 127 MB on 8 B200.  Here every rank maps the other ranks' state
// blocks (CUDA IPC) and ONE cooperative kernel per rank does the whole exchange with ordinary addresses over NVLink.
//
// peer_allreduce_kernel (all-reduce):
//   barrier A   every rank's backward pass is complete (the kernel is stream-ordered behind it)
//   phase 1     rank r sums slice r of all W gradient blocks in rank order 0..W-1 (a reduce-scatter by pulling: cp.async.bulk pieces of every block
//               into shared memory, reduce_slice_tma; or 128-bit loads, reduce_slice) and writes the sum into ALL W blocks in the same loop (the
//               all-gather by pushing; measured on 8 B200: pulling the reduced slices in a second phase 0.47 ms, pushing 0.42, bulk copies 0.37)
//   barrier B   every rank's sums have landed everywhere; nobody reads or writes this rank's block any more
// peer_step_kernel (the exchange fused with the optimiser step):
//   barrier A, phase 1 without the push but with sum g^2 of the slice -> barrier B carries the partial sums to every rank -> clip + TF Adam on the
//   own slice, the updated PARAMETERS written into all W parameter blocks -> barrier C.
// A slice is summed by exactly one rank in a fixed order, so every rank ends up with the same bits and the result does not depend on
// timing.  The barriers are flags in a small exported block per rank: st.release.sys into every peer's block, ld.acquire.sys polls on
// the own one, epochs instead of resets; a watchdog traps instead of hanging.  grid.sync between the local phases.
#pragma once
#include <cooperative_groups.h>
#include "gemm_tcgen05.cuh"      // mbarrier helpers
#include <cuda_runtime.h>
#include <stdint.h>

namespace peer {

constexpr int MAXW = 16;
struct Comm {                      // one per rank, exported; zero-initialised once
    unsigned flag[4][MAXW];        // flag[phase][source rank] = epoch of the last arrival
    double sq[MAXW][8];            // fused optimiser form, written by rank q with barrier B: [0] sum g^2 over its slice, [1] the Wemb part of it,
                                   // [2..4] aux slots (slice square norm, loss, sum(mask)) from the rank that owns them
    double acc[2];                 // this rank's accumulators of [0], [1] (atomics of its own CTAs)
};

struct Args {
    float* g[MAXW];                // gradient blocks of all ranks (own: local pointer)
    Comm* comm[MAXW];              // comm blocks of all ranks
    int rank, world;
    unsigned epoch;
    size_t n4;                     // whole float4 elements in a block; the n - 4 n4 < 4 trailing floats belong to the last rank's slice
    size_t n;                      // floats in a block
    size_t slice4;                 // float4 elements per slice (the last slice may be shorter)
    // fused optimiser form only
    float* p[MAXW];                // parameter blocks of all ranks
    float *m, *v;                  // this rank's Adam slots
    size_t P;                      // parameters (the block holds P + 8 floats: aux slots behind them)
    size_t wemb_lo, wemb_hi;       // float range of the Wemb gradient inside the block
    int use_slice_norm, normalize;
    int tma;                       // phase 1 on the bulk-copy engine (reduce_slice_tma)
    float clip, lr_t, b1, b2, eps;
    float* gnorm_out;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer data is read exactly once per exchange: bypass L1 (no stale lines from the previous exchange), 128 bits per load
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of this GPU have finished their stores -> tell every rank (self included) -> wait for every rank -> release the CTAs
// npayload > 0: doubles handed to every rank's sq[own rank][..] slot together with the arrival (ordered before it by the release)
__device__ __forceinline__ void cross_barrier(const Args& a, int phase, cooperative_groups::grid_group& grid, const double* payload = nullptr, int npayload = 0) {
    __threadfence_system();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)a.world) {
        const int q = threadIdx.x;
        for (int k = 0; k < npayload; ++k) reinterpret_cast<volatile double*>(a.comm[q]->sq[a.rank])[k] = payload[k];
        st_release_sys(&a.comm[q]->flag[phase][a.rank], a.epoch);
        const unsigned* mine = &a.comm[a.rank]->flag[phase][q];
        long long t0 = clock64();
        for (unsigned spins = 1; (int)(ld_acquire_sys(mine) - a.epoch) < 0; ++spins) {
            if ((spins & 1023u) == 0 && clock64() - t0 > 20000000000LL) { printf("s2vt: peer barrier %d timed out waiting for rank %d (rank %d, epoch %u)\n", phase, q, a.rank, a.epoch); __trap(); }
        }
    }
    grid.sync();
}

// phase 1: rank r sums slice r of all W blocks in rank order.  U independent elements per thread and round so that U * (W - 1) >= 4 peer loads are in
// flight per thread whatever the world size.  SQ: also accumulate sum g^2 (all parameters / the Wemb range) of the reduced slice.
template <int WT, int U, bool SQ, bool PUSH>
__device__ __forceinline__ void reduce_slice(const Args& a, size_t tid, size_t nthr, double& sq_all, double& sq_wemb) {
    const int W = a.world, r = a.rank;
    const size_t lo = (size_t)r * a.slice4, hi = lo + a.slice4 < a.n4 ? lo + a.slice4 : a.n4;
    float4* mine = reinterpret_cast<float4*>(a.g[r]);
    const size_t n4p = a.P / 4;
    for (size_t i0 = lo + tid; i0 < hi; i0 += (size_t)U * nthr) {
        float4 v[U][WT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + (size_t)u * nthr;
#pragma unroll
            for (int q = 0; q < WT; ++q)
                if (q < W && i < hi) v[u][q] = q == r ? mine[i] : ld_peer(reinterpret_cast<const float4*>(a.g[q]) + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + (size_t)u * nthr;
            if (i < hi) {
                float4 s = v[u][0];
#pragma unroll
                for (int q = 1; q < WT; ++q)
                    if (q < W) { s.x += v[u][q].x; s.y += v[u][q].y; s.z += v[u][q].z; s.w += v[u][q].w; }
                mine[i] = s;
                if (PUSH) {
#pragma unroll
                    for (int q = 0; q < WT; ++q)
                        if (q < W && q != r) reinterpret_cast<float4*>(a.g[q])[i] = s;
                }
                if (SQ) {
                    const float c[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const size_t j = 4 * i + k;
                        if (i < n4p || j < a.P) {
                            const double d = (double)c[k] * c[k];
                            sq_all += d;
                            if (j >= a.wemb_lo && j < a.wemb_hi) sq_wemb += d;
                        }
                    }
                }
            }
        }
    }
    if (r == W - 1 && tid < a.n - 4 * a.n4) {      // the < 4 trailing floats of the block
        const size_t i = 4 * a.n4 + tid;
        float s = 0.f;
        for (int q = 0; q < W; ++q) s += q == r ? a.g[r][i] : ld_peer1(a.g[q] + i);
        a.g[r][i] = s;
        if (PUSH) for (int q = 0; q < W; ++q) if (q != r) a.g[q][i] = s;
        if (SQ && i < a.P) { sq_all += (double)s * s; if (i >= a.wemb_lo && i < a.wemb_hi) sq_wemb += (double)s * s; }
    }
}
template <bool SQ, bool PUSH>
__device__ __forceinline__ void reduce_slice_any(const Args& a, size_t tid, size_t nthr, double& sq_all, double& sq_wemb) {
    if (a.world <= 2) reduce_slice<2, 4, SQ, PUSH>(a, tid, nthr, sq_all, sq_wemb);
    else if (a.world <= 4) reduce_slice<4, 2, SQ, PUSH>(a, tid, nthr, sq_all, sq_wemb);
    else if (a.world <= 8) reduce_slice<8, 1, SQ, PUSH>(a, tid, nthr, sq_all, sq_wemb);
    else reduce_slice<16, 1, SQ, PUSH>(a, tid, nthr, sq_all, sq_wemb);
}

// phase 2: rank r fetches the slices of the other ranks from their owners' blocks `blk[q]` (the first `nfl` floats of a block count)
__device__ __forceinline__ void gather_slices(const Args& a, float* const* blk, size_t nfl, size_t tid, size_t nthr) {
    const int W = a.world, r = a.rank;
    const size_t n4 = nfl / 4;
    float4* mine = reinterpret_cast<float4*>(blk[r]);
    for (int dq = 1; dq < W; ++dq) {
        const int q = (r + dq) % W;         // every rank starts with another owner: the loads spread over the links
        const size_t lo = (size_t)q * a.slice4, hi = lo + a.slice4 < n4 ? lo + a.slice4 : n4;
        const float4* src = reinterpret_cast<const float4*>(blk[q]);
        size_t i = lo + tid;
        for (; i + 3 * nthr < hi; i += 4 * nthr) {
            const float4 v0 = ld_peer(src + i), v1 = ld_peer(src + i + nthr), v2 = ld_peer(src + i + 2 * nthr), v3 = ld_peer(src + i + 3 * nthr);
            mine[i] = v0; mine[i + nthr] = v1; mine[i + 2 * nthr] = v2; mine[i + 3 * nthr] = v3;
        }
        for (; i < hi; i += nthr) mine[i] = ld_peer(src + i);
        if (q == W - 1 && tid < nfl - 4 * n4) blk[r][4 * n4 + tid] = ld_peer1(blk[q] + 4 * n4 + tid);      // trailing floats: the last rank's
    }
}

// ---- phase 1 on the bulk-copy engine (W <= 8).  SM-issued 128-bit loads / stores move ~530 GB/s per direction over NVLink 5 (measured, 2 and 8 GPUs);
// cp.async.bulk moves whole 8 KB pieces per instruction.  A CTA walks the chunks b, b + grid, ... of slice r: one elected thread requests chunk c of all W
// blocks (W bulk loads completing on one mbarrier, two stages deep), the 512 threads add the W pieces from shared memory in rank order (one float4 each),
// and the sum leaves through a shared staging piece: W bulk stores (PUSH: into every rank's block) or a plain store into the own block.
constexpr int TMA_CH = 2048;                                   // floats per chunk and block: 8 KB
constexpr int TMA_WMAX = 8;
constexpr int TMA_SMEM = 2 * TMA_WMAX * TMA_CH * 4 + TMA_CH * 4 + 64 + 128;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
                 "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes) : "memory");
}

template <bool SQ, bool PUSH>
__device__ __forceinline__ void reduce_slice_tma(const Args& a, unsigned char* smem, double& sq_all, double& sq_wemb) {
    const int W = a.world, r = a.rank;
    float* buf = reinterpret_cast<float*>(smem);                              // [2][TMA_WMAX][TMA_CH]
    float* outp = buf + 2 * TMA_WMAX * TMA_CH;                                // [TMA_CH]
    uint64_t* full = reinterpret_cast<uint64_t*>(outp + TMA_CH);              // [2]
    const size_t lo = (size_t)r * a.slice4 * 4, hi4 = (size_t)r * a.slice4 + a.slice4 < a.n4 ? (size_t)r * a.slice4 + a.slice4 : a.n4, hi = hi4 * 4;   // floats
    const size_t nchunks = hi > lo ? (hi - lo + TMA_CH - 1) / TMA_CH : 0;
    if (threadIdx.x == 0) {
        tc::mbar_init(full, 1); tc::mbar_init(full + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");      // the peers' blocks were written through the generic proxy (ordered by barrier A)
    }
    __syncthreads();
    auto request = [&](size_t c, int st) {                     // thread 0
        const size_t f0 = lo + c * TMA_CH;
        const uint32_t bytes = (uint32_t)((hi - f0 < (size_t)TMA_CH ? hi - f0 : (size_t)TMA_CH) * 4);
        tc::mbar_expect_tx(full + st, bytes * (uint32_t)W);
        for (int q = 0; q < W; ++q) bulk_load(buf + ((size_t)st * TMA_WMAX + q) * TMA_CH, a.g[q] + f0, bytes, full + st);
    };
    size_t k = 0;
    if (threadIdx.x == 0) {
        if (blockIdx.x < nchunks) request(blockIdx.x, 0);
        if (blockIdx.x + (size_t)gridDim.x < nchunks) request(blockIdx.x + gridDim.x, 1);
    }
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++k) {
        const int st = (int)(k & 1);
        tc::mbar_wait(full + st, (uint32_t)((k >> 1) & 1));
        const size_t f0 = lo + c * TMA_CH;
        const size_t nfl = hi - f0 < (size_t)TMA_CH ? hi - f0 : (size_t)TMA_CH;
        const size_t t4 = threadIdx.x;                          // TMA_CH / 4 == blockDim.x: one float4 per thread
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool live = 4 * t4 < nfl;
        if (live) {
            s = reinterpret_cast<const float4*>(buf + ((size_t)st * TMA_WMAX) * TMA_CH)[t4];
            for (int q = 1; q < W; ++q) {
                const float4 v = reinterpret_cast<const float4*>(buf + ((size_t)st * TMA_WMAX + q) * TMA_CH)[t4];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            if (SQ) {
                const float cc[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const size_t e = f0 + 4 * t4 + j;
                    if (e < a.P) { const double d = (double)cc[j] * cc[j]; sq_all += d; if (e >= a.wemb_lo && e < a.wemb_hi) sq_wemb += d; }
                }
            }
        }
        if (PUSH) {
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous chunk's stores have read the staging piece
            __syncthreads();
            if (live) reinterpret_cast<float4*>(outp)[t4] = s;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();                                     // staging piece complete, stage st free
            if (threadIdx.x == 0) {
                for (int dq = 0; dq < W; ++dq) { const int q = (r + dq) % W; bulk_store(a.g[q] + f0, outp, (uint32_t)(nfl * 4)); }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            if (live) reinterpret_cast<float4*>(a.g[r] + f0)[t4] = s;
            __syncthreads();                                     // stage st free
        }
        if (threadIdx.x == 0 && c + 2 * (size_t)gridDim.x < nchunks) request(c + 2 * (size_t)gridDim.x, st);
    }
    if (PUSH && threadIdx.x == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");             // all stores of this CTA are complete (performed), not merely read
        asm volatile("fence.proxy.async;" ::: "memory");                      // ... and ordered before the generic-proxy release of barrier B
    }
    if (r == W - 1 && blockIdx.x == 0 && threadIdx.x < a.n - 4 * a.n4) {      // the < 4 trailing floats of the block
        const size_t i = 4 * a.n4 + threadIdx.x;
        float sum = 0.f;
        for (int q = 0; q < W; ++q) sum += q == r ? a.g[r][i] : ld_peer1(a.g[q] + i);
        a.g[r][i] = sum;
        if (PUSH) for (int q = 0; q < W; ++q) if (q != r) a.g[q][i] = sum;
        if (SQ && i < a.P) { sq_all += (double)sum * sum; if (i >= a.wemb_lo && i < a.wemb_hi) sq_wemb += (double)sum * sum; }
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned char* peer_smem() {
    extern __shared__ unsigned char peer_smem_raw[];
    return reinterpret_cast<unsigned char*>(((uintptr_t)peer_smem_raw + 127) & ~(uintptr_t)127);
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(Args a) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    double d0 = 0, d1 = 0;
    cross_barrier(a, 0, grid);
    if (a.tma) reduce_slice_tma<false, true>(a, peer_smem(), d0, d1);
    else reduce_slice_any<false, true>(a, tid, nthr, d0, d1);
    cross_barrier(a, 1, grid);
}

// The exchange fused with the optimiser step (ZeRO-1 style): reduce-scatter -> global norm from per-rank partial sums -> clip + TF Adam on the own slice
// only (1 / W of the 0.9 GB that the full Adam pass moves) -> the updated PARAMETERS are pushed into every rank's block instead of the gradients.  Same arithmetic per element as
// adam_kernel.  The Adam slots m / v of a rank are current in its own slice only (s2vt_peer_gather_state before saving them).
__global__ void __launch_bounds__(512) peer_step_kernel(Args a) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ double red[2][16];
    __shared__ double payload[8];
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    const int W = a.world, r = a.rank;
    Comm* my = a.comm[r];
    if (tid == 0) { my->acc[0] = 0.0; my->acc[1] = 0.0; }
    cross_barrier(a, 0, grid);
    double sq_all = 0.0, sq_wemb = 0.0;
    if (a.tma) reduce_slice_tma<true, false>(a, peer_smem(), sq_all, sq_wemb);
    else reduce_slice_any<true, false>(a, tid, nthr, sq_all, sq_wemb);
    {   // CTA reduction, one atomic pair per CTA
        for (int o = 16; o > 0; o >>= 1) { sq_all += __shfl_xor_sync(0xffffffffu, sq_all, o); sq_wemb += __shfl_xor_sync(0xffffffffu, sq_wemb, o); }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sq_all; red[1][threadIdx.x >> 5] = sq_wemb; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s0 = 0, s1 = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s0 += red[0][w]; s1 += red[1][w]; }
            atomicAdd(&my->acc[0], s0); atomicAdd(&my->acc[1], s1);
        }
    }
    __threadfence();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        payload[0] = *reinterpret_cast<volatile double*>(&my->acc[0]); payload[1] = *reinterpret_cast<volatile double*>(&my->acc[1]);
        for (int k = 0; k < 3; ++k) payload[2 + k] = r == W - 1 ? (double)*reinterpret_cast<volatile float*>(a.g[r] + a.P + k) : 0.0;
    }
    __syncthreads();
    cross_barrier(a, 1, grid, payload, 5);
    // every rank: the same sums in the same (rank) order
    double s_all = 0.0, s_wemb = 0.0;
    for (int q = 0; q < W; ++q) { s_all += reinterpret_cast<volatile double*>(my->sq[q])[0]; s_wemb += reinterpret_cast<volatile double*>(my->sq[q])[1]; }
    const double aux_slice = reinterpret_cast<volatile double*>(my->sq[W - 1])[2], aux_loss = reinterpret_cast<volatile double*>(my->sq[W - 1])[3],
                 aux_mask = reinterpret_cast<volatile double*>(my->sq[W - 1])[4];
    const float inv = a.normalize ? 1.0f / (float)aux_mask : 1.0f;
    double s = s_all;
    if (a.use_slice_norm) s = s - s_wemb + aux_slice;
    s *= (double)inv * (double)inv;
    const float gn = (float)sqrt(s);
    float scale = 1.0f;
    if (a.clip > 0.f && gn > 0.f) scale = a.clip * fminf(1.0f / gn, 1.0f / a.clip);
    if (a.gnorm_out && tid == 0) { a.gnorm_out[0] = gn; a.gnorm_out[1] = (float)aux_loss * inv; }
    // (The same phase with the parameters leaving through a shared staging piece and W bulk stores per 8 KB chunk was measured: slower, 0.55 vs 0.52 ms alone
    // on 2 GPUs -- two CTA barriers per chunk and one float4 per thread starve the local loads of g / m / v / p.)
    {   // Adam on the own slice
        const size_t n4p = a.P / 4;
        const size_t lo = (size_t)r * a.slice4, hi0 = lo + a.slice4 < a.n4 ? lo + a.slice4 : a.n4, hi = hi0 < n4p ? hi0 : n4p;
        float4* t4 = reinterpret_cast<float4*>(a.p[r]); const float4* g4 = reinterpret_cast<const float4*>(a.g[r]);
        float4* m4 = reinterpret_cast<float4*>(a.m); float4* v4 = reinterpret_cast<float4*>(a.v);
        const float b1 = a.b1, b2 = a.b2, eps = a.eps, lr_t = a.lr_t;
        for (size_t i = lo + tid; i < hi; i += nthr) {
            const float4 gg = g4[i]; float4 mm = m4[i], vv = v4[i], tt = t4[i];
            float gi;
            gi = gg.x * inv * scale; mm.x = b1 * mm.x + (1.f - b1) * gi; vv.x = b2 * vv.x + (1.f - b2) * gi * gi; tt.x -= lr_t * mm.x / (sqrtf(vv.x) + eps);
            gi = gg.y * inv * scale; mm.y = b1 * mm.y + (1.f - b1) * gi; vv.y = b2 * vv.y + (1.f - b2) * gi * gi; tt.y -= lr_t * mm.y / (sqrtf(vv.y) + eps);
            gi = gg.z * inv * scale; mm.z = b1 * mm.z + (1.f - b1) * gi; vv.z = b2 * vv.z + (1.f - b2) * gi * gi; tt.z -= lr_t * mm.z / (sqrtf(vv.z) + eps);
            gi = gg.w * inv * scale; mm.w = b1 * mm.w + (1.f - b1) * gi; vv.w = b2 * vv.w + (1.f - b2) * gi * gi; tt.w -= lr_t * mm.w / (sqrtf(vv.w) + eps);
            m4[i] = mm; v4[i] = vv; t4[i] = tt;
            for (int q = 0; q < W; ++q)
                if (q != r) reinterpret_cast<float4*>(a.p[q])[i] = tt;          // the all-gather of the parameters, pushed from the registers
        }
        if (r == W - 1 && tid < a.P - 4 * n4p) {      // the < 4 trailing parameters
            const size_t i = 4 * n4p + tid;
            const float gi = a.g[r][i] * inv * scale;
            const float mi = b1 * a.m[i] + (1.f - b1) * gi, vi = b2 * a.v[i] + (1.f - b2) * gi * gi;
            a.m[i] = mi; a.v[i] = vi;
            const float ti = a.p[r][i] - lr_t * mi / (sqrtf(vi) + eps);
            for (int q = 0; q < W; ++q) a.p[q][i] = ti;
        }
    }
    cross_barrier(a, 2, grid);      // every rank's slice has landed in this rank's parameter block
}

// m / v of all ranks' slices into this rank's blocks (before a checkpoint is written): blk = the adam_m or adam_v blocks of all ranks
__global__ void __launch_bounds__(512) peer_gather_kernel(Args a, int phase0) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    cross_barrier(a, phase0, grid);
    gather_slices(a, a.p, a.P, tid, nthr);
    cross_barrier(a, phase0 + 1, grid);
}

}  // namespace peer
