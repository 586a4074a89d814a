// GEMM building blocks: C[M,N] = A[M,K] . B[N,K]^T  (both operands K-major, N and K multiples of 128,
// M guarded), accumulators handed to an epilogue functor through a shared-memory tile.
//
// Mainloops:
//   * Mainloop<bf16,...>  : mma.sync.m16n8k16 bf16 -> fp32, cp.async 3-stage pipeline (warp-level tensor path; the
//                           recurrent-step kernels use it and it is the fallback/checker for the tcgen05 GEMM).
//   * Mainloop<float,...> : SIMT fp32 FMA (the 1e-5 "fp32 mode" of BASELINE.json; also the on-device checker).
//   * gemm_tcgen05.cuh    : tcgen05.mma + TMA + TMEM for the large batched GEMMs.
// Epilogues (all read the fp32 tile Cs[BM][LDC] that the mainloop leaves in shared memory):
//   EpiStore, EpiGradStore, EpiLstmFwd, EpiLstmBwd.
#pragma once
#include "common.cuh"

// -----------------------------------------------------------------------------------------------------------------
// tile configurations
// -----------------------------------------------------------------------------------------------------------------
template <int BM_, int BN_, int WM_, int WN_, int BK_ = 32, int ST_ = 3>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_;
    static constexpr int NWARPS = (BM / WM) * (BN / WN);
    static constexpr int NTHREADS = NWARPS * 32;
    static constexpr int LDC = BN + 8;  // fp32 epilogue tile row stride
    static constexpr int BK = BK_, STAGES = ST_, LDS = BK + 8;
    static constexpr size_t PIPE_BYTES = (size_t)STAGES * (BM + BN) * LDS * 2;
    static constexpr size_t SIMT_BYTES = (size_t)16 * (BM + 4 + BN + 4) * 4;
    static constexpr size_t EPI_BYTES = (size_t)BM * LDC * 4;
    static constexpr size_t SMEM_BYTES =
        (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) > SIMT_BYTES ? (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) : SIMT_BYTES;
};
typedef TileCfg<128, 128, 64, 32> CfgBig;    // batched GEMMs: 8 warps, warp tile 64x32
typedef TileCfg<64, 32, 16, 32, 128, 3> CfgStep;   // recurrent-step GEMMs: 4 warps, many CTAs for small M, long K blocks (few barriers)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
template <typename T16> __device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma_16816<bf16>(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma_16816<f16>(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T, class Cfg> struct Mainloop;

// ---- 16-bit tensor-core mainloop (mma.sync; bf16 x bf16 or fp16 x fp16) -----------------------------------------
template <typename T16, class Cfg>
struct Mainloop16 {
    __device__ static void run(const T16* __restrict__ A_, int lda, const T16* __restrict__ B_, int ldb, int M, int K, int m0, int n0,
                               unsigned char* smem, float* Cs) {
        const bf16* A = reinterpret_cast<const bf16*>(A_); const bf16* B = reinterpret_cast<const bf16*>(B_);   // 16-bit payloads: the copy code is type-blind
        constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, LDS = Cfg::LDS, ST = Cfg::STAGES;
        constexpr int MT = WM / 16, NT = WN / 8;
        static_assert(NT % 2 == 0, "warp tile N must cover pairs of n8 tiles");
        bf16* sA = reinterpret_cast<bf16*>(smem);
        bf16* sB = sA + ST * BM * LDS;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int wm = warp / (BN / WN), wn = warp % (BN / WN);
        const int KT = K / Cfg::BK;

        auto load_stage = [&](int s, int kt) {
            bf16* a = sA + s * BM * LDS;
            bf16* b = sB + s * BN * LDS;
            constexpr int CH = Cfg::BK / 8;   // 16-byte chunks per tile row
            for (int c = tid; c < BM * CH; c += Cfg::NTHREADS) {
                int r = c / CH, ch = c % CH, gr = m0 + r;
                bool ok = gr < M;
                cp_async16(a + r * LDS + ch * 8, A + (size_t)(ok ? gr : 0) * lda + kt * Cfg::BK + ch * 8, ok ? 16 : 0);
            }
            for (int c = tid; c < BN * CH; c += Cfg::NTHREADS) {
                int r = c / CH, ch = c % CH;
                cp_async16(b + r * LDS + ch * 8, B + (size_t)(n0 + r) * ldb + kt * Cfg::BK + ch * 8, 16);
            }
        };

        float acc[MT][NT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

#pragma unroll
        for (int s = 0; s < ST - 1; ++s) {
            if (s < KT) load_stage(s, s);
            cp_async_commit();
        }
        for (int kt = 0; kt < KT; ++kt) {
            cp_async_wait<ST - 2>();
            __syncthreads();
            if (kt + ST - 1 < KT) load_stage((kt + ST - 1) % ST, kt + ST - 1);
            cp_async_commit();
            const bf16* a = sA + (kt % ST) * BM * LDS;
            const bf16* b = sB + (kt % ST) * BN * LDS;
#pragma unroll
            for (int kk = 0; kk < Cfg::BK; kk += 16) {
                uint32_t af[MT][4];
#pragma unroll
                for (int i = 0; i < MT; ++i)
                    ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3], a + (wm * WM + i * 16 + (lane & 15)) * LDS + kk + (lane >> 4) * 8);
#pragma unroll
                for (int j = 0; j < NT; j += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(b0, b1, b2, b3, b + (wn * WN + j * 8 + (lane & 7) + (lane >> 4) * 8) * LDS + kk + ((lane >> 3) & 1) * 8);
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        mma_16816<T16>(acc[i][j], af[i], b0, b1);
                        mma_16816<T16>(acc[i][j + 1], af[i], b2, b3);
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // every warp is done with the pipeline buffers; Cs aliases them
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                int r = wm * WM + i * 16 + (lane >> 2), c = wn * WN + j * 8 + (lane & 3) * 2;
                *reinterpret_cast<float2*>(Cs + r * Cfg::LDC + c) = make_float2(acc[i][j][0], acc[i][j][1]);
                *reinterpret_cast<float2*>(Cs + (r + 8) * Cfg::LDC + c) = make_float2(acc[i][j][2], acc[i][j][3]);
            }
        __syncthreads();
    }
};

template <class Cfg> struct Mainloop<bf16, Cfg> : Mainloop16<bf16, Cfg> {};
template <class Cfg> struct Mainloop<f16, Cfg> : Mainloop16<f16, Cfg> {};

// ---- fp32 SIMT mainloop ----------------------------------------------------------------------------------------
template <class Cfg>
struct Mainloop<float, Cfg> {
    __device__ static void run(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M, int K, int m0, int n0,
                               unsigned char* smem, float* Cs) {
        constexpr int BM = Cfg::BM, BN = Cfg::BN, NTH = Cfg::NTHREADS, BK = 16;
        constexpr int TGN = BN / 4, TGM = NTH / TGN, TM = BM / TGM;
        static_assert(TGM * TM == BM && TGN * 4 == BN, "bad SIMT thread grid");
        constexpr int LDA = BM + 4, LDB = BN + 4;
        float* As = reinterpret_cast<float*>(smem);  // [BK][LDA]
        float* Bs = As + BK * LDA;                   // [BK][LDB]
        const int tid = threadIdx.x, tx = tid % TGN, ty = tid / TGN;
        float acc[TM][4];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
        for (int k0 = 0; k0 < K; k0 += BK) {
            for (int c = tid; c < BM * (BK / 4); c += NTH) {
                int r = c % BM, ch = c / BM, gr = m0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < M) v = *reinterpret_cast<const float4*>(A + (size_t)gr * lda + k0 + ch * 4);
                As[(ch * 4 + 0) * LDA + r] = v.x; As[(ch * 4 + 1) * LDA + r] = v.y;
                As[(ch * 4 + 2) * LDA + r] = v.z; As[(ch * 4 + 3) * LDA + r] = v.w;
            }
            for (int c = tid; c < BN * (BK / 4); c += NTH) {
                int r = c % BN, ch = c / BN;
                float4 v = *reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * ldb + k0 + ch * 4);
                Bs[(ch * 4 + 0) * LDB + r] = v.x; Bs[(ch * 4 + 1) * LDB + r] = v.y;
                Bs[(ch * 4 + 2) * LDB + r] = v.z; Bs[(ch * 4 + 3) * LDB + r] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float4 b = *reinterpret_cast<const float4*>(Bs + k * LDB + tx * 4);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    float a = As[k * LDA + ty + i * TGM];
                    acc[i][0] = fmaf(a, b.x, acc[i][0]); acc[i][1] = fmaf(a, b.y, acc[i][1]);
                    acc[i][2] = fmaf(a, b.z, acc[i][2]); acc[i][3] = fmaf(a, b.w, acc[i][3]);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
            *reinterpret_cast<float4*>(Cs + (ty + i * TGM) * Cfg::LDC + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __syncthreads();
    }
};

// -----------------------------------------------------------------------------------------------------------------
// epilogues
// -----------------------------------------------------------------------------------------------------------------
// 8 consecutive values -> compute dtype, vector stores (dst 16-byte aligned)
__device__ __forceinline__ void store8(float* dst, const float* v) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* dst, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void store8(f16* dst, const float* v) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// out = acc (+ bias[col]) (+ old out if accumulate); written as fp32 and/or compute dtype.
template <typename T>
struct EpiStore {
    static constexpr bool kDirect = false;
    struct Params {
        float* outF; T* outT; int ldo; const float* bias; int M; int accumulate;
    };
    template <class Cfg>
    __device__ static void apply(const Params& p, const float* Cs, int m0, int n0) {
        for (int idx = threadIdx.x; idx < Cfg::BM * Cfg::BN / 4; idx += Cfg::NTHREADS) {
            int r = idx / (Cfg::BN / 4), c = (idx % (Cfg::BN / 4)) * 4, gr = m0 + r, gc = n0 + c;
            if (gr >= p.M) continue;
            float4 v = *reinterpret_cast<const float4*>(Cs + r * Cfg::LDC + c);
            if (p.bias) { float4 b = *reinterpret_cast<const float4*>(p.bias + gc); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
            size_t o = (size_t)gr * p.ldo + gc;
            if (p.outF) {
                if (p.accumulate) { float4 w = *reinterpret_cast<const float4*>(p.outF + o); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
                *reinterpret_cast<float4*>(p.outF + o) = v;
            }
            if (p.outT) store_gates4(p.outT + o, v);      // 4 consecutive values: one 8-byte (16-bit types) / 16-byte (fp32) store
        }
    }
};

// Vocabulary projection fused with the decoder's word choice (rollouts): the logits never reach HBM.  Each CTA reduces its
// tile of (logit + bias [+ Gumbel noise for sampling rows]) to one (value, index) candidate per row; the next step's cell
// kernel takes the arg-max over the tiles' candidates (EpiLstmFwd::prefetch).  Noise and tie-breaking are those of
// sample_rows_kernel (same Philox stream, lowest index wins), so the chosen words are identical.
template <typename T>
struct EpiLogitsPick {
    static constexpr bool kDirect = false;
    struct Params {
        int M, V; const float* bias; int n_sample; unsigned long long seed; uint32_t step, row_base;
        float* pick_val; int* pick_idx; int pick_ld;     // [M, pick_ld] candidates, column = tile index
    };
    static constexpr int kEpiWarps = 16;   // tcgen05 kernel: 16 epilogue warps (the Philox / log work needs the lanes)
    // apply_on: the functor run by threads [t0, t0 + NTH) of the CTA only, synchronised through named barrier 1 (the persistent
    // sampling chain keeps its TMA / MMA warps busy with the next step's cell GEMM meanwhile); apply: by the whole CTA.
    template <class Cfg>
    __device__ static void apply(const Params& p, const float* Cs, int m0, int n0) { apply_on<Cfg, 0, Cfg::NTHREADS>(p, Cs, m0, n0); }
    template <class Cfg, int T0, int NTH>
    __device__ static void apply_on(const Params& p, const float* Cs, int m0, int n0) {
        constexpr int PARTS = 4, PW = Cfg::BN / PARTS;          // each row's tile columns are scanned by 4 threads
        __shared__ float part_v[PARTS][Cfg::BM];
        __shared__ int part_i[PARTS][Cfg::BM];
        const int tile = n0 / Cfg::BN;
        const int tid = (int)threadIdx.x - T0;
        for (int u = tid; u < Cfg::BM * PARTS; u += NTH) {
            const int r = u % Cfg::BM, part = u / Cfg::BM, gr = m0 + r;   // consecutive threads -> consecutive rows (conflict-free float4 reads)
            ArgVal best; best.v = -INFINITY; best.i = 0x7fffffff;
            if (gr < p.M) {
                const bool sample = gr < p.n_sample;
                const float* row = Cs + r * Cfg::LDC;
                const int cb = part * PW;
                float cut = -INFINITY;
                if (sample) {   // words > 19.5 below the local maximum cannot win (bounded noise, see sample_rows_kernel)
                    float mx = -INFINITY;
                    for (int c = cb; c < cb + PW; c += 4) {
                        float4 x = *reinterpret_cast<const float4*>(row + c), b = *reinterpret_cast<const float4*>(p.bias + n0 + c);
                        if (n0 + c + 0 < p.V) mx = fmaxf(mx, x.x + b.x);
                        if (n0 + c + 1 < p.V) mx = fmaxf(mx, x.y + b.y);
                        if (n0 + c + 2 < p.V) mx = fmaxf(mx, x.z + b.z);
                        if (n0 + c + 3 < p.V) mx = fmaxf(mx, x.w + b.w);
                    }
                    cut = mx - 22.0f;
                }
                // up to 16 words per iteration: the Philox blocks are independent, so their 10-round multiply chains overlap (one block at
                // a time left the warp waiting on its own dependent instructions -- issue slots 36 % busy in the ncu capture).  Noise is
                // still skipped for groups that cannot win; adding it to hopeless words of a live group cannot change the winner.
                constexpr int CH = PW % 16 == 0 ? 16 : (PW % 8 == 0 ? 8 : 4), NG = CH / 4;
                for (int c = cb; c < cb + PW; c += CH) {
                    float e[CH];
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const float4 x = *reinterpret_cast<const float4*>(row + c + 4 * g), b = *reinterpret_cast<const float4*>(p.bias + n0 + c + 4 * g);
                        e[4 * g] = x.x + b.x; e[4 * g + 1] = x.y + b.y; e[4 * g + 2] = x.z + b.z; e[4 * g + 3] = x.w + b.w;
                    }
                    if (sample) {
                        float gm = e[0];
#pragma unroll
                        for (int k = 1; k < CH; ++k) gm = fmaxf(gm, e[k]);
                        if (gm < cut) continue;
                        uint4 o[NG];
#pragma unroll
                        for (int g = 0; g < NG; ++g)
                            o[g] = philox4x32_10((uint32_t)((n0 + c + 4 * g) >> 2), p.step, p.row_base + (uint32_t)gr, S2VT_STREAM_SAMPLE, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            e[4 * g] += gumbel_fast(u32_to_uniform(o[g].x)); e[4 * g + 1] += gumbel_fast(u32_to_uniform(o[g].y));
                            e[4 * g + 2] += gumbel_fast(u32_to_uniform(o[g].z)); e[4 * g + 3] += gumbel_fast(u32_to_uniform(o[g].w));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < CH; ++k)
                        if (n0 + c + k < p.V) { ArgVal cand; cand.v = e[k]; cand.i = n0 + c + k; best = argmax_op(best, cand); }
                }
            }
            part_v[part][r] = best.v; part_i[part][r] = best.i;
        }
        if constexpr (T0 == 0 && NTH == Cfg::NTHREADS) __syncthreads();
        else asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory");
        for (int r = tid; r < Cfg::BM; r += NTH) {
            const int gr = m0 + r;
            if (gr >= p.M) continue;
            ArgVal best; best.v = part_v[0][r]; best.i = part_i[0][r];
#pragma unroll
            for (int q = 1; q < PARTS; ++q) { ArgVal c; c.v = part_v[q][r]; c.i = part_i[q][r]; best = argmax_op(best, c); }
            p.pick_val[(size_t)gr * p.pick_ld + tile] = best.v;
            p.pick_idx[(size_t)gr * p.pick_ld + tile] = best.i;
        }
    }
};

// Vocabulary projection fused with the beam step's word ranking (final_beam_search.py:213-217: softmax + top_k): the
// logits never reach HBM.  A tile row is scanned in PARTS column parts; each part emits its TOPK_MAX best words in
// (value descending, index ascending) order -- tf.nn.top_k's order -- and its soft-max statistics (max, sum exp(l - max)).
// beam_step_kernel merges the parts: a row's top-k words are among the parts' top-k, and lse = log sum_p s_p e^{m_p}.
#define TOPK_MAX 8
template <typename T>
struct EpiLogitsTopK {
    static constexpr bool kDirect = false;
    static constexpr int PARTS = 4;
    struct Params {
        int M, V; const float* bias;
        float* cand_val; int* cand_idx;   // [M, nparts, TOPK_MAX]
        float2* stat;                     // [M, nparts]
        int nparts;                       // Vp / (BN / PARTS)
    };
    static constexpr int kEpiWarps = 16;
    template <class Cfg>
    __device__ static void apply(const Params& p, const float* Cs, int m0, int n0) {
        constexpr int PW = Cfg::BN / PARTS;
        static_assert(PW % 4 == 0, "part width");
        for (int u = threadIdx.x; u < Cfg::BM * PARTS; u += Cfg::NTHREADS) {
            const int r = u % Cfg::BM, part = u / Cfg::BM, gr = m0 + r;   // consecutive threads -> consecutive rows
            if (gr >= p.M) continue;
            const int c0 = n0 + part * PW;                                // first word of the part
            const float* row = Cs + r * Cfg::LDC + part * PW;
            const float* bias = p.bias + c0;
            float mx = -INFINITY;
            for (int c = 0; c < PW; c += 4) {
                const float4 x = *reinterpret_cast<const float4*>(row + c), b = *reinterpret_cast<const float4*>(bias + c);
                if (c0 + c + 0 < p.V) mx = fmaxf(mx, x.x + b.x);
                if (c0 + c + 1 < p.V) mx = fmaxf(mx, x.y + b.y);
                if (c0 + c + 2 < p.V) mx = fmaxf(mx, x.z + b.z);
                if (c0 + c + 3 < p.V) mx = fmaxf(mx, x.w + b.w);
            }
            float tv[TOPK_MAX]; int ti[TOPK_MAX];
#pragma unroll
            for (int q = 0; q < TOPK_MAX; ++q) { tv[q] = -INFINITY; ti[q] = 0x7fffffff; }
            float se = 0.f;
            for (int c = 0; c < PW; c += 4) {
                const float4 x = *reinterpret_cast<const float4*>(row + c), b = *reinterpret_cast<const float4*>(bias + c);
                const float e[4] = {x.x + b.x, x.y + b.y, x.z + b.z, x.w + b.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (c0 + c + k >= p.V) continue;
                    se += expf(e[k] - mx);
                    if (e[k] > tv[TOPK_MAX - 1]) {          // ascending scan + strict compares: the lower index stays ahead on ties
                        tv[TOPK_MAX - 1] = e[k]; ti[TOPK_MAX - 1] = c0 + c + k;
#pragma unroll
                        for (int q = TOPK_MAX - 1; q > 0; --q)
                            if (tv[q] > tv[q - 1]) {
                                const float fv = tv[q]; tv[q] = tv[q - 1]; tv[q - 1] = fv;
                                const int fi = ti[q]; ti[q] = ti[q - 1]; ti[q - 1] = fi;
                            }
                    }
                }
            }
            const size_t o = (size_t)gr * p.nparts + c0 / PW;
            float4* cv = reinterpret_cast<float4*>(p.cand_val + o * TOPK_MAX);
            int4* ci = reinterpret_cast<int4*>(p.cand_idx + o * TOPK_MAX);
            cv[0] = make_float4(tv[0], tv[1], tv[2], tv[3]); cv[1] = make_float4(tv[4], tv[5], tv[6], tv[7]);
            ci[0] = make_int4(ti[0], ti[1], ti[2], ti[3]); ci[1] = make_int4(ti[4], ti[5], ti[6], ti[7]);
            p.stat[o] = make_float2(mx, se);
        }
    }
};

// Weight-gradient store into the fp32 TF-layout gradient block:  grad[(row0 + r) * ldg + colmap(c)] += scale * acc.
// gate_h > 0 : columns are in packed gate order (c = 4u+g) and map to the TF order g*gate_h + u (u < gate_h).
// gate_h == 0: identity columns, valid while c < ncols.   Rows valid while r < nrows.
struct EpiGradStore {
    static constexpr bool kDirect = false;
    struct Params {
        float* grad; int ldg; int nrows; int ncols; int gate_h; float scale;
        int atomic;      // != 0: several CTAs (split-K) add into the same elements -> red.global.add instead of a plain read-modify-write
    };
    template <class Cfg>
    __device__ static void apply(const Params& p, const float* Cs, int m0, int n0) {
        // read-modify-write of 128 x BN gradient elements per CTA: 8 independent elements per thread and round, all loads of a round issued before the first
        // store (one element at a time the loop was a chain of dependent L2 round trips: ~85 us per launch, more than the mainloop of most weight gradients)
        constexpr int PER = 8, UN = Cfg::BN / 4;
        if (p.atomic) {      // split-K: fire-and-forget reductions, nothing to wait for (the plain loop measured 10 us faster than the batched form)
            for (int idx = threadIdx.x; idx < Cfg::BM * Cfg::BN; idx += Cfg::NTHREADS) {
                const int r = idx / Cfg::BN, rem = idx % Cfg::BN, gr = m0 + r;
                if (gr >= p.nrows) continue;
                if (p.gate_h > 0) {
                    const int g = rem / UN, ul = rem % UN, u = n0 / 4 + ul;
                    if (u < p.gate_h) atomicAdd(p.grad + (size_t)gr * p.ldg + (size_t)g * p.gate_h + u, p.scale * Cs[r * Cfg::LDC + ul * 4 + g]);
                } else if (n0 + rem < p.ncols) {
                    atomicAdd(p.grad + (size_t)gr * p.ldg + n0 + rem, p.scale * Cs[r * Cfg::LDC + rem]);
                }
            }
            return;
        }
        for (int base = threadIdx.x; base < Cfg::BM * Cfg::BN; base += Cfg::NTHREADS * PER) {
            float* dst[PER]; float v[PER], old[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = base + i * Cfg::NTHREADS;
                dst[i] = nullptr;
                if (idx < Cfg::BM * Cfg::BN) {
                    const int r = idx / Cfg::BN, rem = idx % Cfg::BN, gr = m0 + r;
                    if (p.gate_h > 0) {      // thread -> (row, gate, unit): consecutive threads take consecutive units of one gate (coalesced stores)
                        const int g = rem / UN, ul = rem % UN, u = n0 / 4 + ul;
                        if (gr < p.nrows && u < p.gate_h) { dst[i] = p.grad + (size_t)gr * p.ldg + (size_t)g * p.gate_h + u; v[i] = p.scale * Cs[r * Cfg::LDC + ul * 4 + g]; }
                    } else {
                        const int gc = n0 + rem;
                        if (gr < p.nrows && gc < p.ncols) { dst[i] = p.grad + (size_t)gr * p.ldg + gc; v[i] = p.scale * Cs[r * Cfg::LDC + rem]; }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PER; ++i) if (dst[i]) old[i] = *dst[i];
#pragma unroll
            for (int i = 0; i < PER; ++i) if (dst[i]) *dst[i] = old[i] + v[i];
        }
    }
};

// Fused BasicLSTMCell forward (Q7).  acc = h_prev . Wh (packed gate columns 4u+g);
//   pre = acc + bias + add0[row0(row)] + add1[tok[row]] ; i,j,f,o -> c' = c*sig(f+1) + sig(i)*tanh(j) ; h' = tanh(c')*sig(o)
template <typename T>
struct EpiLstmFwd {
    // Direct form (tcgen05 path): a thread owns one accumulator row and a chunk of 32 packed gate columns = 8 whole
    // units, so the cell is evaluated straight from registers; `prefetch` issues every global load of the chunk before
    // the accumulator is ready (the epilogue warps are idle during the mainloop).
    static constexpr bool kDirect = true;
    struct Pre { float4 add[8]; float4 c[2]; };
    struct Params;
    __device__ static void prefetch(const Params& p, int gr, int gc, Pre& pre);
    __device__ static void direct(const Params& p, int gr, int gc, const float* v, const Pre& pre);
    struct Params {
        int M; int Hp;
        const float* bias;                 // [4Hp] packed
        const float* add0; int add0_mod;   // [*, 4Hp] input-side pre-activations; row = add0_mod > 0 ? row % add0_mod : row
        const T* add1; const int* tok;     // embedding table [V, 4Hp] gathered by tok[row] (nullable), in the forward operand type
        const float* pick_val; const int* pick_idx; int pick_ld, pick_nt;   // alternative to tok: per-tile candidates of EpiLogitsPick
        int* tok_out; int* ids_out; int ids_ld, ids_col;                    // ...resolved here; CTA column 0 records the word
        const float* c_prev; float* c_out; // [M, Hp]
        T* h_out;                          // [M, Hp] compute dtype (next step's A operand / batched GEMM operand)
        float* h_outF;                     // optional fp32 copy
        T* gates_out;                      // optional [M, 4Hp] saved activations (si, tj, sf, so) for BPTT, compute dtype
        T* hdrop_out;                      // optional dropout-applied output (DropoutWrapper, Q2)
        unsigned long long seed; uint32_t stream; uint32_t step; uint32_t row_base; float keep;
    };
    template <class Cfg>
    __device__ static void apply(const Params& p, const float* Cs, int m0, int n0) {
        constexpr int UN = Cfg::BN / 4;
        const int G = 4 * p.Hp;
        for (int idx = threadIdx.x; idx < Cfg::BM * UN; idx += Cfg::NTHREADS) {
            int r = idx / UN, ul = idx % UN, gr = m0 + r;
            if (gr >= p.M) continue;
            int gc = n0 + ul * 4, u = gc >> 2;
            float4 v = *reinterpret_cast<const float4*>(Cs + r * Cfg::LDC + ul * 4);
            float4 b = *reinterpret_cast<const float4*>(p.bias + gc);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            if (p.add0) {
                int ar = p.add0_mod > 0 ? gr % p.add0_mod : gr;
                float4 a = *reinterpret_cast<const float4*>(p.add0 + (size_t)ar * G + gc);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            if (p.add1) {
                float4 a = load_gates4(p.add1 + (size_t)resolve_token(p, gr, gc) * G + gc);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            float si = sigm<T>(v.x), tj = tanh_<T>(v.y), sf = sigm<T>(v.z + 1.0f), so = sigm<T>(v.w);
            size_t o = (size_t)gr * p.Hp + u;
            float c = p.c_prev[o] * sf + si * tj;
            float h = tanh_<T>(c) * so;
            p.c_out[o] = c;
            p.h_out[o] = from_f32<T>(h);
            if (p.h_outF) p.h_outF[o] = h;
            if (p.gates_out) store_gates4(p.gates_out + (size_t)gr * G + gc, make_float4(si, tj, sf, so));
            if (p.hdrop_out) {
                float m = p.keep < 1.0f ? dropout_mult(p.seed, p.stream, p.row_base + gr, p.step, u, p.keep) : 1.0f;
                p.hdrop_out[o] = from_f32<T>(h * m);
            }
        }
    }
};

// 32 consecutive table values (one thread's 8 units x 4 gates) as 8 float4: 128 B of fp32 or 64 B of fp16 (four 16-byte loads in flight)
__device__ __forceinline__ void load_row32(const float* a, float4* x) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = a ? reinterpret_cast<const float4*>(a)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void load_row32(const f16* a, float4* x) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = a ? reinterpret_cast<const uint4*>(a)[j] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2* hp = reinterpret_cast<const __half2*>(&u[j]);
        const float2 a0 = __half22float2(hp[0]), a1 = __half22float2(hp[1]), a2 = __half22float2(hp[2]), a3 = __half22float2(hp[3]);
        x[2 * j] = make_float4(a0.x, a0.y, a1.x, a1.y); x[2 * j + 1] = make_float4(a2.x, a2.y, a3.x, a3.y);
    }
}
__device__ __forceinline__ void load_row32(const bf16* a, float4* x) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = a ? load_gates4(a + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <class P>
__device__ __forceinline__ int resolve_token(const P& p, int gr, int gc) {
    if (!p.pick_val) return p.tok[gr];
    const float* pv = p.pick_val + (size_t)gr * p.pick_ld;
    const int nt = p.pick_nt;
    float bv = pv[0]; int bt = 0;                            // first maximum = lowest word index
    if ((p.pick_ld & 1) == 0) {
        // 16 candidates per round, every load issued before the first compare: one L2 round trip per round instead of one per 32-byte
        // sector (this scan sits on the token -> Etab row -> gate pre-activation chain that nothing can overlap)
        const float2* pv2 = reinterpret_cast<const float2*>(pv);
        for (int t0 = 0; t0 < nt; t0 += 16) {
            float2 c[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c[q] = t0 + 2 * q < nt ? pv2[(t0 >> 1) + q] : make_float2(-INFINITY, -INFINITY);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (c[q].x > bv) { bv = c[q].x; bt = t0 + 2 * q; }
                if (t0 + 2 * q + 1 < nt && c[q].y > bv) { bv = c[q].y; bt = t0 + 2 * q + 1; }
            }
        }
    } else {
        for (int t = 1; t < nt; ++t) { float v = pv[t]; if (v > bv) { bv = v; bt = t; } }
    }
    const int w = p.pick_idx[(size_t)gr * p.pick_ld + bt];
    if (gc == 0) { if (p.tok_out) p.tok_out[gr] = w; if (p.ids_out) p.ids_out[(size_t)gr * p.ids_ld + p.ids_col] = w; }
    return w;
}
template <typename T>
__device__ __forceinline__ void EpiLstmFwd<T>::prefetch(const Params& p, int gr, int gc, Pre& pre) {
    if (gr >= p.M) return;
    const size_t G = 4 * (size_t)p.Hp;
    const float4* b = reinterpret_cast<const float4*>(p.bias + gc);
    const float4* a0 = p.add0 ? reinterpret_cast<const float4*>(p.add0 + (size_t)(p.add0_mod > 0 ? gr % p.add0_mod : gr) * G + gc) : nullptr;
    const float4* c = reinterpret_cast<const float4*>(p.c_prev + (size_t)gr * p.Hp + (gc >> 2));
    float4 x0[8], x1[8];
    // everything that does not depend on the word first: these loads are in flight while the candidates are resolved
#pragma unroll
    for (int j = 0; j < 8; ++j) { pre.add[j] = b[j]; x0[j] = a0 ? a0[j] : make_float4(0.f, 0.f, 0.f, 0.f); }
    pre.c[0] = c[0]; pre.c[1] = c[1];
    const T* a1 = p.add1 ? p.add1 + (size_t)resolve_token(p, gr, gc) * G + gc : nullptr;
    load_row32(a1, x1);
#pragma unroll
    for (int j = 0; j < 8; ++j) pre.add[j] = f4add(pre.add[j], f4add(x0[j], x1[j]));
}
template <typename T>
__device__ __forceinline__ void EpiLstmFwd<T>::direct(const Params& p, int gr, int gc, const float* v, const Pre& pre) {
    if (gr >= p.M) return;
    const size_t G = 4 * (size_t)p.Hp;
    const int u0 = gc >> 2;
    const float cp[8] = {pre.c[0].x, pre.c[0].y, pre.c[0].z, pre.c[0].w, pre.c[1].x, pre.c[1].y, pre.c[1].z, pre.c[1].w};
    float cn[8], hn[8];
    T* gsave = p.gates_out ? p.gates_out + (size_t)gr * G + gc : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float si = sigm<T>(v[4 * j] + pre.add[j].x), tj = tanh_<T>(v[4 * j + 1] + pre.add[j].y);
        float sf = sigm<T>(v[4 * j + 2] + pre.add[j].z + 1.0f), so = sigm<T>(v[4 * j + 3] + pre.add[j].w);
        cn[j] = cp[j] * sf + si * tj;
        hn[j] = tanh_<T>(cn[j]) * so;
        if (gsave) store_gates4(gsave + 4 * j, make_float4(si, tj, sf, so));      // stored as produced: nothing but c', h' stays live
    }
    const size_t o = (size_t)gr * p.Hp + u0;
    store8(p.c_out + o, cn);
    store8(p.h_out + o, hn);
    if (p.h_outF) store8(p.h_outF + o, hn);
    if (p.hdrop_out) {
        if (p.keep < 1.0f) {
            float4 m0 = dropout_mult4(p.seed, p.stream, p.row_base + gr, p.step, u0, p.keep);
            float4 m1 = dropout_mult4(p.seed, p.stream, p.row_base + gr, p.step, u0 + 4, p.keep);
            hn[0] *= m0.x; hn[1] *= m0.y; hn[2] *= m0.z; hn[3] *= m0.w; hn[4] *= m1.x; hn[5] *= m1.y; hn[6] *= m1.z; hn[7] *= m1.w;
        }
        store8(p.hdrop_out + o, hn);
    }
}

// BasicLSTMCell backward for one time step.  dh = dh_rec + dh_ext * dropout ; produces the pre-activation gate
// gradients (packed order) and dc for the previous step.
struct LstmBwdArgs {
    int M; int Hp;
    const float* dh_ext;     // [M, Hp] nullable: gradient arriving from the layer above at this step
    const void* gates;       // [M, 4Hp] (si, tj, sf, so) in the compute dtype
    const float* c_prev;     // [M, Hp]
    const float* c_new;      // [M, Hp]
    float* dc;               // [M, Hp] in: dc from step t+1, out: dc for step t-1
    unsigned long long seed; uint32_t stream; uint32_t step; uint32_t row_base; float keep;  // dropout on dh_ext (keep>=1: none)
};
// T: type of the gate GRADIENTS written; F: type the forward pass saved the gate ACTIVATIONS in (Fwd<T>::type in the S2VT engine)
template <typename T, typename F = T>
__device__ __forceinline__ void lstm_bwd_unit(const LstmBwdArgs& p, T* dg_out, int gr, int u, float dh_rec) {
    const int G = 4 * p.Hp;
    size_t o = (size_t)gr * p.Hp + u;
    float dh = dh_rec;
    if (p.dh_ext) {
        float m = p.keep < 1.0f ? dropout_mult(p.seed, p.stream, p.row_base + gr, p.step, u, p.keep) : 1.0f;
        dh += p.dh_ext[o] * m;
    }
    float4 g = load_gates4(reinterpret_cast<const F*>(p.gates) + (size_t)gr * G + 4 * u);
    float si = g.x, tj = g.y, sf = g.z, so = g.w;
    float tc = tanh_<T>(p.c_new[o]);
    float d_o = dh * tc;
    float dc = p.dc[o] + dh * so * (1.0f - tc * tc);
    float di = dc * tj, dj = dc * si, df = dc * p.c_prev[o];
    p.dc[o] = dc * sf;
    T* d = dg_out + (size_t)gr * G + 4 * u;
    d[0] = from_f32<T>(di * si * (1.0f - si));
    d[1] = from_f32<T>(dj * (1.0f - tj * tj));
    d[2] = from_f32<T>(df * sf * (1.0f - sf));
    d[3] = from_f32<T>(d_o * so * (1.0f - so));
}
// GEMM form: acc[row, u] = dG(t+1)[row, :] . Wh[u, :]   (N dimension = hidden units)
template <typename T, typename F = T>
struct EpiLstmBwd {
    // Direct form: a thread owns one row and a chunk of 8 hidden units of dh_rec (the accumulator columns are units).
    static constexpr bool kDirect = true;
    static constexpr int kUnitsPerChunk = 8;
    struct Pre { GateRaw<F> g[8]; float4 cn[2], cp[2], dc[2], dh[2]; };   // gates stay packed until used (register pressure of the split-K epilogue)
    struct Params {
        LstmBwdArgs a; T* dg_out;
    };
    // part 0: everything; 1: what the split-K epilogue fetches before the mainloop ends (gates, dc, dh_ext); 2: the rest (c_new, c_prev),
    // fetched after the partial tiles have been sent so that fewer registers are live while the 32-column partial is held
    __device__ static void prefetch(const Params& p, int gr, int u0, Pre& pre, int part = 0) {
        const LstmBwdArgs& a = p.a;
        if (gr >= a.M) return;
        const size_t o = (size_t)gr * a.Hp + u0;
        if (part != 2) {
            const F* g = reinterpret_cast<const F*>(a.gates) + (size_t)gr * 4 * a.Hp + 4 * u0;
#pragma unroll
            for (int j = 0; j < 8; ++j) load_gates_raw(g + 4 * j, pre.g[j]);
            const float4* dc = reinterpret_cast<const float4*>(a.dc + o); pre.dc[0] = dc[0]; pre.dc[1] = dc[1];
            if (a.dh_ext) { const float4* dh = reinterpret_cast<const float4*>(a.dh_ext + o); pre.dh[0] = dh[0]; pre.dh[1] = dh[1]; }
            else { pre.dh[0] = pre.dh[1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
        if (part != 1) {
            const float4* cn = reinterpret_cast<const float4*>(a.c_new + o); pre.cn[0] = cn[0]; pre.cn[1] = cn[1];
            const float4* cp = reinterpret_cast<const float4*>(a.c_prev + o); pre.cp[0] = cp[0]; pre.cp[1] = cp[1];
        }
    }
    // v: dh_rec of the 8 units
    __device__ static void direct(const Params& p, int gr, int u0, const float* v, const Pre& pre) {
        const LstmBwdArgs& a = p.a;
        if (gr >= a.M) return;
        const size_t o = (size_t)gr * a.Hp + u0;
        const float cn[8] = {pre.cn[0].x, pre.cn[0].y, pre.cn[0].z, pre.cn[0].w, pre.cn[1].x, pre.cn[1].y, pre.cn[1].z, pre.cn[1].w};
        const float cp[8] = {pre.cp[0].x, pre.cp[0].y, pre.cp[0].z, pre.cp[0].w, pre.cp[1].x, pre.cp[1].y, pre.cp[1].z, pre.cp[1].w};
        const float dcn[8] = {pre.dc[0].x, pre.dc[0].y, pre.dc[0].z, pre.dc[0].w, pre.dc[1].x, pre.dc[1].y, pre.dc[1].z, pre.dc[1].w};
        float dhx[8] = {pre.dh[0].x, pre.dh[0].y, pre.dh[0].z, pre.dh[0].w, pre.dh[1].x, pre.dh[1].y, pre.dh[1].z, pre.dh[1].w};
        if (a.dh_ext && a.keep < 1.0f) {
            float4 m0 = dropout_mult4(a.seed, a.stream, a.row_base + gr, a.step, u0, a.keep);
            float4 m1 = dropout_mult4(a.seed, a.stream, a.row_base + gr, a.step, u0 + 4, a.keep);
            dhx[0] *= m0.x; dhx[1] *= m0.y; dhx[2] *= m0.z; dhx[3] *= m0.w; dhx[4] *= m1.x; dhx[5] *= m1.y; dhx[6] *= m1.z; dhx[7] *= m1.w;
        }
        float dcp[8];
        T* d = p.dg_out + (size_t)gr * 4 * a.Hp + 4 * u0;
#pragma unroll
        for (int j2 = 0; j2 < 8; j2 += 2) {      // two units = 8 gate gradients = one 16-byte (bf16) / two 16-byte (fp32) stores, issued as produced
            float dg[8];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = j2 + jj;
                const float4 gj = unpack_gates(pre.g[j]);
                float si = gj.x, tj = gj.y, sf = gj.z, so = gj.w;
                float dh = v[j] + dhx[j];
                float tc = tanh_<T>(cn[j]);
                float d_o = dh * tc;
                float dc = dcn[j] + dh * so * (1.0f - tc * tc);
                dcp[j] = dc * sf;
                dg[4 * jj] = dc * tj * si * (1.0f - si);
                dg[4 * jj + 1] = dc * si * (1.0f - tj * tj);
                dg[4 * jj + 2] = dc * cp[j] * sf * (1.0f - sf);
                dg[4 * jj + 3] = d_o * so * (1.0f - so);
            }
            store8(d + 4 * j2, dg);
        }
        store8(a.dc + o, dcp);
    }
    template <class Cfg>
    __device__ static void apply(const Params& p, const float* Cs, int m0, int n0) {
        for (int idx = threadIdx.x; idx < Cfg::BM * Cfg::BN; idx += Cfg::NTHREADS) {
            int r = idx / Cfg::BN, c = idx % Cfg::BN, gr = m0 + r;
            if (gr >= p.a.M) continue;
            lstm_bwd_unit<T, F>(p.a, p.dg_out, gr, n0 + c, Cs[r * Cfg::LDC + c]);
        }
    }
};

template <typename T, class Cfg, class Epi>
__global__ void __launch_bounds__(Cfg::NTHREADS) gemm_kernel(const T* __restrict__ A, int lda, const T* __restrict__ B, int ldb, int M, int K,
                                                             typename Epi::Params ep) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int m0 = blockIdx.y * Cfg::BM, n0 = blockIdx.x * Cfg::BN;
    float* Cs = reinterpret_cast<float*>(smem);
    Mainloop<T, Cfg>::run(A, lda, B, ldb, M, K, m0, n0, smem, Cs);
    Epi::template apply<Cfg>(ep, Cs, m0, n0);
}

// Host launcher.  N, K are padded sizes (multiples of 128).
template <typename T, class Cfg, class Epi>
inline cudaError_t launch_gemm(cudaStream_t st, const T* A, int lda, const T* B, int ldb, int M, int N, int K, const typename Epi::Params& ep) {
    if (M <= 0) return cudaSuccess;
    auto kern = gemm_kernel<T, Cfg, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(N / Cfg::BN, (M + Cfg::BM - 1) / Cfg::BM);
    kern<<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>>(A, lda, B, ldb, M, K, ep);
    return cudaGetLastError();
}
