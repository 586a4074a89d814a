// Persistent tcgen05 GEMM for the wide products with plain stores (G1x, G2x, logits, Etab, dout1; fp32 and / or 16-bit output):  C[M,N] = A[M,K] . B[N,K]^T + bias.
//
// With K = 512 .. 1024 a 128 x 256 tile is 8 .. 16 K-blocks of mainloop followed by 128 KB of fp32 stores; in gemm_tc_kernel the two
// run one after the other in every CTA (the staged epilogue re-uses the pipeline's shared memory), and at 1 CTA per SM nothing else
// covers the stores: G2x of a training pass writes 603 MB, the logits 447 MB.  Here one CTA per SM walks a static list of tiles:
//   warp 0      TMA producer, runs ahead across tile boundaries (4-stage ring, never drained)
//   warp 1      tcgen05.mma issuer, alternating between TWO TMEM accumulators (2 x 256 columns)
//   warps 2..9  epilogue: TMEM -> registers (+ bias) -> a padded 32 x 33 shared-memory patch per warp (transposes the "thread owns a row"
//               layout of tcgen05.ld into "warp owns a row") -> coalesced 128-byte global stores,
// so the stores of tile i run under the mainloop of tile i+1.  Same accumulation order as gemm_tc_kernel (bit-identical results).
#pragma once
#include "gemm_tcgen05.cuh"

namespace tc {

constexpr int PS_BN = 256, PS_STAGES = 4, PS_EPI_WARPS = 8, PS_THREADS = 64 + 32 * PS_EPI_WARPS;
constexpr int PS_STAGE_BYTES = BM * BK * 2 + PS_BN * BK * 2;               // 48 KB
constexpr int PS_PATCH = 32 * 33 * 4;                                      // per-warp transpose patch
constexpr int PS_SMEM = PS_STAGES * PS_STAGE_BYTES + PS_EPI_WARPS * PS_PATCH + 256 + 1024;

template <typename T16>
__global__ void __launch_bounds__(PS_THREADS) gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                     int M, int N, int K, uint32_t fmt, float* __restrict__ out, T16* __restrict__ out16, int ldo,
                                                                     const float* __restrict__ bias) {
    const uint32_t IDESC = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PS_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24)) & ~fmt;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* patches = reinterpret_cast<float*>(smem + PS_STAGES * PS_STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + PS_STAGES * PS_STAGE_BYTES + PS_EPI_WARPS * PS_PATCH);
    uint64_t* empty = full + PS_STAGES;
    uint64_t* acc_full = empty + PS_STAGES;                               // [2]
    uint64_t* acc_empty = acc_full + 2;                                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntn = N / PS_BN, ntm = (M + BM - 1) / BM, ntiles = ntn * ntm;
    const int KBL = K / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < PS_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, PS_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
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
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int m0 = (t / ntn) * BM, n0 = (t % ntn) * PS_BN;        // column tiles of one row block are neighbours in the list: A rows hit in L2
            for (int i = 0; i < KBL; ++i, ++g) {
                const int st = g % PS_STAGES;
                if (g >= PS_STAGES) mbar_wait(empty + st, ((g / PS_STAGES) - 1) & 1);
                if (leader) {
                    unsigned char* a = smem + st * PS_STAGE_BYTES;
                    mbar_expect_tx(full + st, (uint32_t)PS_STAGE_BYTES);
                    tma_load_2d_raw(a, &mapA, full + st, i * BK, m0);
                    tma_load_2d_raw(a + BM * BK * 2, &mapB, full + st, i * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const bool leader = elect_one();
        int g = 0, j = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            if (j >= 2) mbar_wait(acc_empty + buf, ((j >> 1) - 1) & 1);    // the epilogue has drained this accumulator (tile j - 2)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int i = 0; i < KBL; ++i, ++g) {
                const int st = g % PS_STAGES;
                mbar_wait(full + st, (g / PS_STAGES) & 1);
                const uint32_t a = smem_u32(smem + st * PS_STAGE_BYTES);
                const uint64_t adesc = make_desc(a), bdesc = make_desc(a + BM * BK * 2);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) mma_bf16(tmem_base + (uint32_t)(buf * PS_BN), adesc + 2 * k, bdesc + 2 * k, IDESC, i > 0 || k != 0);
                    mma_commit(empty + st);
                }
            }
            if (leader) mma_commit(acc_full + buf);
        }
        __syncwarp();
    } else {
        // epilogue warp e: TMEM lane quarter q = warp % 4, column half (e / 4): chunks of 32 columns [128 (e/4), +128)
        const int e = warp - 2, q = warp & 3, half = e >> 2;
        float* patch = patches + e * (32 * 33);
        int j = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            const int m0 = (t / ntn) * BM, n0 = (t % ntn) * PS_BN;
            mbar_wait(acc_full + buf, (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * PS_BN + half * 128);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                tmem_ld32(trow + (uint32_t)(32 * c), v);
                const int gc = n0 + half * 128 + 32 * c;
                __syncwarp();                                             // the previous chunk's reads of the patch are done
#pragma unroll
                for (int k = 0; k < 32; ++k) patch[lane * 33 + k] = v[k];   // lane = row of the quarter: bank (lane + k) % 32, conflict-free
                __syncwarp();
                const float bv = bias ? bias[gc + lane] : 0.f;            // lane = column now
                const int rbase = m0 + q * 32;
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int gr = rbase + r;
                    if (gr < M) {
                        const float x = patch[r * 33 + lane] + bv;
                        if (out) out[(size_t)gr * ldo + gc + lane] = x;                             // 32 lanes: one 128-byte line
                        if (out16) out16[(size_t)gr * ldo + gc + lane] = from_f32<T16>(x);          // ... or one 64-byte half line
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_empty + buf)) : "memory");
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// fp32 and / or 16-bit output (ldo elements per row), optional bias[N]; N % 256 == 0, K % 64 == 0.
template <typename T16>
inline cudaError_t launch_persist(MapCache& cache, cudaStream_t st, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, float* out, T16* out16, int ldo,
                                  const float* bias, bool pdl, uint32_t fmt) {
    if (M <= 0) return cudaSuccess;
    if (cache.size() > 32768) cache.clear();
    const CUtensorMap* ma = get_map(cache, A, M, K, lda, BM);
    const CUtensorMap* mb = get_map(cache, B, N, K, ldb, PS_BN);
    if (!ma || !mb) return cudaErrorInvalidValue;
    static int sms = 0;
    if (!sms) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_persist_kernel<T16>, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int ntiles = (N / PS_BN) * ((M + BM - 1) / BM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles < sms ? ntiles : sms, 1, 1);
    cfg.blockDim = dim3(PS_THREADS);
    cfg.dynamicSmemBytes = PS_SMEM;
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
    return cudaLaunchKernelEx(&cfg, gemm_tc_persist_kernel<T16>, *ma, *mb, M, N, K, fmt, out, out16, ldo, bias);
}

}  // namespace tc
