// Element-wise / row kernels of the S2VT path (HBM-bound work: vectorised, coalesced, no tensor cores).
#pragma once
#include "common.cuh"

// ---- weight packing: fp32 TF-layout master -> compute-dtype padded K-major copies --------------------------------
// src: fp32 [R, C] (row stride lds) starting at row r0 of a TF variable.
// dst: T, zero-initialised by the caller.  gate_h > 0: the C dimension is 4*gate_h in TF gate order (g*H+u) and is
// re-ordered to the packed order 4u+g.  transpose: dst[c'][r] (ld = ldd) else dst[r][c'].
template <typename T>
__global__ void pack_matrix_kernel(const float* __restrict__ src, int lds, int R, int C, T* __restrict__ dst, int ldd, int gate_h, int transpose) {
    __shared__ float tile[32][33];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * lds + c] : 0.f;
    }
    __syncthreads();
    if (transpose) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            int c = c0 + i, r = r0 + threadIdx.x;   // consecutive threads -> consecutive r (contiguous in dst)
            if (r < R && c < C) {
                int cp = gate_h > 0 ? 4 * (c % gate_h) + c / gate_h : c;
                dst[(size_t)cp * ldd + r] = from_f32<T>(tile[threadIdx.x][i]);
            }
        }
    } else {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            int r = r0 + i, c = c0 + threadIdx.x;
            if (r < R && c < C) {
                int cp = gate_h > 0 ? 4 * (c % gate_h) + c / gate_h : c;
                dst[(size_t)r * ldd + cp] = from_f32<T>(tile[i][threadIdx.x]);
            }
        }
    }
}

// fp32 vector with optional gate re-ordering into a zero-initialised padded fp32 vector.
__global__ void pack_vector_kernel(const float* __restrict__ src, int C, float* __restrict__ dst, int gate_h) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) dst[gate_h > 0 ? 4 * (c % gate_h) + c / gate_h : c] = src[c];
}

// video fp32 [B, Tv, D] (batch-major, the reference feed layout) -> compute dtype, time-major [Tv*B, Dp], zero padded.
// src_index (nullable) maps output video slot b -> source video row (row de-duplication / replication).
template <typename T>
__global__ void convert_video_kernel(const float* __restrict__ video, const int* __restrict__ src_index, int B, int Tv, int D, int Dp, T* __restrict__ out) {
    int row = blockIdx.x;  // t*B + b
    int t = row / B, b = row % B;
    int sb = src_index ? src_index[b] : b;
    const float* s = video + ((size_t)sb * Tv + t) * D;
    T* d = out + (size_t)row * Dp;
    for (int c = threadIdx.x; c < Dp; c += blockDim.x) d[c] = from_f32<T>(c < D ? s[c] : 0.f);
}

// Generic tiled transpose of a compute-dtype matrix: dst[c][r] = src[r][c], r < R, c < C (dst padded region untouched); TI != TO
// converts on the way (fp16 activations -> bf16 for the K-major checker mainloops).
template <typename TI, typename TO = TI>
__global__ void transpose_kernel(const TI* __restrict__ src, int lds, int R, int C, TO* __restrict__ dst, int ldd) {
    __shared__ TO tile[32][34];
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = from_f32<TO>(to_f32(src[(size_t)r * lds + c]));
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) dst[(size_t)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// element-wise conversion between the two 16-bit types, 8 elements (one 128-bit access) per thread and iteration; n % 8 == 0 and both
// pointers 16-byte aligned (padded activation matrices)
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ src, size_t n, TO* __restrict__ dst) {
    const size_t n8 = n / 8, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 u = reinterpret_cast<const uint4*>(src)[i];
        const TI* x = reinterpret_cast<const TI*>(&u);
        uint4 o;
        TO* y = reinterpret_cast<TO*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = from_f32<TO>(to_f32(x[k]));
        reinterpret_cast<uint4*>(dst)[i] = o;
    }
}

// fp32 -> compute dtype copy with optional dropout (used for the dropout-applied LSTM1 output of the training graph):
//   out[t*N + n, u] = h[(t)*B + n % B, u] * mult(global row n, step t, u)
template <typename T>
__global__ void expand_dropout_kernel(const T* __restrict__ h, int B, int N, int Hp, int H, T* __restrict__ out, unsigned long long seed, uint32_t stream,
                                      uint32_t row_base, float keep) {
    int row = blockIdx.x;  // t*N + n
    int t = row / N, n = row % N;
    const T* s = h + ((size_t)t * B + (n % B)) * Hp;
    T* d = out + (size_t)row * Hp;
    for (int u4 = threadIdx.x * 4; u4 < Hp; u4 += blockDim.x * 4) {
        float4 m = keep < 1.0f ? dropout_mult4(seed, stream, row_base + n, t, u4, keep) : make_float4(1.f, 1.f, 1.f, 1.f);
        d[u4] = from_f32<T>(to_f32(s[u4]) * m.x); d[u4 + 1] = from_f32<T>(to_f32(s[u4 + 1]) * m.y);
        d[u4 + 2] = from_f32<T>(to_f32(s[u4 + 2]) * m.z); d[u4 + 3] = from_f32<T>(to_f32(s[u4 + 3]) * m.w);
    }
}

// Gradient w.r.t. LSTM1's (shared per video) output: dh1_ext[t*B + b, u] = sum_k dout1[t*N + k*B + b, u] * mult(n, t, u)
__global__ void reduce_dropout_kernel(const float* __restrict__ dout1, int B, int N, int Hp, float* __restrict__ dh1, unsigned long long seed, uint32_t stream,
                                      uint32_t row_base, float keep) {
    int row = blockIdx.x;  // t*B + b
    int t = row / B, b = row % B;
    for (int u4 = threadIdx.x * 4; u4 < Hp; u4 += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = b; n < N; n += B) {
            float4 m = keep < 1.0f ? dropout_mult4(seed, stream, row_base + n, t, u4, keep) : make_float4(1.f, 1.f, 1.f, 1.f);
            float4 v = *reinterpret_cast<const float4*>(dout1 + ((size_t)t * N + n) * Hp + u4);
            acc.x += v.x * m.x; acc.y += v.y * m.y; acc.z += v.z * m.z; acc.w += v.w * m.w;
        }
        *reinterpret_cast<float4*>(dh1 + (size_t)row * Hp + u4) = acc;
    }
}

// ---- vocabulary row kernels (one CTA per row, 128-bit loads, row kept in registers) -------------------------------
#define ROW_THREADS 256
#define ROW_MAXV4 12  // supports V <= 256*12*4 = 12288

// Rollout step: rows < n_sample use Gumbel-max categorical sampling (tf.multinomial [lib]), the rest argmax (tf.argmax).
// Writes the token for the next step's embedding gather and the id matrix column.
__global__ void __launch_bounds__(ROW_THREADS) sample_rows_kernel(const float* __restrict__ logits, int ld, int V, int n_sample, unsigned long long seed,
                                                                   uint32_t step, uint32_t row_base, int* __restrict__ tok_out,
                                                                   int* __restrict__ ids_out, int ids_ld) {
    __shared__ ArgVal red[32];
    int row = blockIdx.x;
    const float4* lr = reinterpret_cast<const float4*>(logits + (size_t)row * ld);
    bool sample = row < n_sample;
    uint32_t grow = row_base + (uint32_t)row;
    // Sampling rows: with the 23-bit uniforms of u32_to_uniform the Gumbel noise lies in [-2.8, 16.7], so a word more than
    // 19.5 below the row maximum can never win the arg-max.  Such words are skipped (exactly, not approximately: the result
    // equals the full arg-max over the same Philox stream); one cheap max pass finds the cut.
    __shared__ float redf[32];
    float cut = -INFINITY;
    if (sample) {
        float mx = -INFINITY;
        for (int v4 = threadIdx.x; v4 * 4 < V; v4 += ROW_THREADS) {
            float4 x = lr[v4];
            int v = v4 * 4;
            if (v + 0 < V) mx = fmaxf(mx, x.x);
            if (v + 1 < V) mx = fmaxf(mx, x.y);
            if (v + 2 < V) mx = fmaxf(mx, x.z);
            if (v + 3 < V) mx = fmaxf(mx, x.w);
        }
        mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, redf);
        cut = mx - 22.0f;
    }
    ArgVal best; best.v = -INFINITY; best.i = 0x7fffffff;
    for (int v4 = threadIdx.x; v4 * 4 < V; v4 += ROW_THREADS) {
        float4 x = lr[v4];
        float e[4] = {x.x, x.y, x.z, x.w};
        if (sample) {
            if (fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)) < cut) continue;
            uint4 o = philox4x32_10((uint32_t)v4, step, grow, S2VT_STREAM_SAMPLE, (uint32_t)seed, (uint32_t)(seed >> 32));
            uint32_t rr[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] += gumbel_fast(u32_to_uniform(rr[k]));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int v = v4 * 4 + k;
            if (v < V) { ArgVal c; c.v = e[k]; c.i = v; best = argmax_op(best, c); }
        }
    }
    best = block_argmax(best, red);
    if (threadIdx.x == 0) {
        tok_out[row] = best.i;
        if (ids_out) ids_out[(size_t)row * ids_ld + step] = best.i;
    }
}

// Training row kernel: log-softmax statistics + fused softmax backward.
//   logp[row]   = logit[w] - lse                       (log-prob of the target word, un-masked)
//   sumlsm[row] = sum_v (logit_v - lse)                (needed by the label-smoothed CE, Q3)
//   dlogits[row, v] = ca[row] * softmax_v - cb[row] * [v == w] - cc[row]     (R4; XE: ca = c, cb = (1-ls) c, cc = c ls / V)
// Padded columns (v >= V) of dlogits are written as zero.  ca == nullptr skips the backward.
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) softmax_rows_kernel(const float* __restrict__ logits, int ld, int V, int Vp, const int* __restrict__ target,
                                                                    const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ cc,
                                                                    float* __restrict__ logp, float* __restrict__ sumlsm, T* __restrict__ dlogits,
                                                                    float* __restrict__ logp_bm, int N, int Tc) {
    __shared__ float red[32];
    int row = blockIdx.x;
    const float4* lr = reinterpret_cast<const float4*>(logits + (size_t)row * ld);
    float4 x[ROW_MAXV4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < ROW_MAXV4; ++i) {
        int v4 = threadIdx.x + i * ROW_THREADS;
        if (v4 * 4 < Vp) {
            x[i] = lr[v4];
            int v = v4 * 4;
            if (v + 0 < V) mx = fmaxf(mx, x[i].x);
            if (v + 1 < V) mx = fmaxf(mx, x[i].y);
            if (v + 2 < V) mx = fmaxf(mx, x[i].z);
            if (v + 3 < V) mx = fmaxf(mx, x[i].w);
        }
    }
    mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, red);
    float se = 0.f, sl = 0.f;
#pragma unroll
    for (int i = 0; i < ROW_MAXV4; ++i) {
        int v4 = threadIdx.x + i * ROW_THREADS;
        if (v4 * 4 < Vp) {
            int v = v4 * 4;
            if (v + 0 < V) { se += expf(x[i].x - mx); sl += x[i].x; }
            if (v + 1 < V) { se += expf(x[i].y - mx); sl += x[i].y; }
            if (v + 2 < V) { se += expf(x[i].z - mx); sl += x[i].z; }
            if (v + 3 < V) { se += expf(x[i].w - mx); sl += x[i].w; }
        }
    }
    se = block_reduce(se, [](float a, float b) { return a + b; }, red);
    sl = block_reduce(sl, [](float a, float b) { return a + b; }, red);
    float lse = mx + logf(se);
    int w = target[row];
    if (threadIdx.x == 0) {
        float lp = logits[(size_t)row * ld + w] - lse;
        logp[row] = lp;
        if (logp_bm) logp_bm[(size_t)(row % N) * Tc + row / N] = lp;   // batch-major [N, Tc] copy for the caller
        if (sumlsm) sumlsm[row] = sl - (float)V * lse;
    }
    if (ca == nullptr) return;
    float a = ca[row], b = cb[row], c = cc ? cc[row] : 0.f;
    T* dr = dlogits + (size_t)row * Vp;
#pragma unroll
    for (int i = 0; i < ROW_MAXV4; ++i) {
        int v4 = threadIdx.x + i * ROW_THREADS;
        if (v4 * 4 < Vp) {
            int v = v4 * 4;
            float e[4] = {x[i].x, x[i].y, x[i].z, x[i].w}, d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) d[k] = v + k < V ? a * expf(e[k] - lse) - (v + k == w ? b : 0.f) - c : 0.f;
            store_gates4(dr + v, make_float4(d[0], d[1], d[2], d[3]));      // one 8-byte (bf16) / 16-byte (fp32) store instead of four scalar ones
        }
    }
}

// Beam step: per row softmax statistics + top-k (k <= 8) by repeated block argmax.  prob = exp(l - lse) in fp32 (B6).
__global__ void __launch_bounds__(ROW_THREADS) topk_rows_kernel(const float* __restrict__ logits, int ld, int V, int k, int* __restrict__ idx_out,
                                                                 float* __restrict__ logp_out) {
    __shared__ ArgVal red[32];
    __shared__ float redf[32];
    int row = blockIdx.x;
    const float* lr = logits + (size_t)row * ld;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += ROW_THREADS) mx = fmaxf(mx, lr[v]);
    mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, redf);
    float se = 0.f;
    for (int v = threadIdx.x; v < V; v += ROW_THREADS) se += expf(lr[v] - mx);
    se = block_reduce(se, [](float a, float b) { return a + b; }, redf);
    float lse = mx + logf(se);
    int taken[8];
    for (int j = 0; j < k; ++j) {
        ArgVal best; best.v = -INFINITY; best.i = 0x7fffffff;
        for (int v = threadIdx.x; v < V; v += ROW_THREADS) {
            bool skip = false;
            for (int q = 0; q < j; ++q) skip |= (taken[q] == v);
            if (!skip) { ArgVal c; c.v = lr[v]; c.i = v; best = argmax_op(best, c); }
        }
        best = block_argmax(best, red);
        taken[j] = best.i;
        if (threadIdx.x == 0) { idx_out[row * k + j] = best.i; logp_out[row * k + j] = best.v - lse; }
    }
}

// Last decode step of a rollout with fused word choice: arg-max over the per-tile candidates -> ids[:, col].
__global__ void resolve_picks_kernel(const float* __restrict__ pick_val, const int* __restrict__ pick_idx, int ld, int nt, int R, int* __restrict__ ids,
                                     int ids_ld, int col) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* pv = pick_val + (size_t)r * ld;
    float bv = pv[0]; int bt = 0;
    for (int t = 1; t < nt; ++t) { float v = pv[t]; if (v > bv) { bv = v; bt = t; } }
    ids[(size_t)r * ids_ld + col] = pick_idx[(size_t)r * ld + bt];
}

// ---- small utilities --------------------------------------------------------------------------------------------
__global__ void fill_int_kernel(int* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Caption mask (R1, cider_evaluation.py:156-171): 1 through the first <eos>(0), 0 after; also lengths (tokens before <eos>).
__global__ void caption_mask_kernel(const int* __restrict__ ids, int N, int Tc, float* __restrict__ mask, int* __restrict__ len) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int L = Tc, done = 0;
    for (int t = 0; t < Tc; ++t) {
        if (mask) mask[(size_t)n * Tc + t] = done ? 0.f : 1.f;
        if (!done && ids[(size_t)n * Tc + t] == 0) { done = 1; L = t; }
    }
    if (len) len[n] = L;
}

// Time-major teacher-forcing tables from the [N, Tc] caption matrix:
//   prev_tok[i*N + n] = (i == 0) ? <bos>=1 : caption[n, i-1] ;  target[i*N + n] = caption[n, i]
__global__ void caption_tables_kernel(const int* __restrict__ cap, int N, int Tc, int* __restrict__ prev_tok, int* __restrict__ target) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Tc) return;
    int i = idx / N, n = idx % N;
    prev_tok[idx] = i == 0 ? 1 : cap[(size_t)n * Tc + i - 1];
    target[idx] = cap[(size_t)n * Tc + i];
}

// RL / XE per-row softmax-backward coefficients (time-major rows i*N + n).
//  mode 0 (REINFORCE, R4): ca = cb = scale * (r_n - b_n) * mask[n,i] / norm ; cc = 0
//  mode 1 (XE, Q3)       : c = scale * S_i / (N * norm) ; ca = c ; cb = (1-ls) c ; cc = c * ls / V
__global__ void loss_coef_kernel(int mode, const float* __restrict__ mask, const float* __restrict__ rewards, const float* __restrict__ base, int N, int Tc,
                                 const float* __restrict__ norm_p, float scale, float ls, int V, float* __restrict__ ca, float* __restrict__ cb,
                                 float* __restrict__ cc, const float* __restrict__ colsum_global = nullptr, int n_global = 0) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Tc) return;
    const float norm = norm_p[0];
    int i = idx / N, n = idx % N;
    if (mode == 0) {
        float c = scale * (rewards[n] - base[n]) * mask[(size_t)n * Tc + i] / norm;
        ca[idx] = c; cb[idx] = c; cc[idx] = 0.f;
    } else {
        // Q3: step loss = mean_b(CE_b) * sum_b mask[b, i] over the WHOLE batch; a data-parallel caller passes the global column sums
        float S = 0.f;
        if (colsum_global) S = colsum_global[i];
        else for (int m = 0; m < N; ++m) S += mask[(size_t)m * Tc + i];
        float c = scale * S / ((float)(colsum_global ? n_global : N) * norm);
        ca[idx] = c; cb[idx] = (1.f - ls) * c; cc[idx] = c * ls / (float)V;
    }
}

// Scalar losses from the per-row statistics.  out[0] = RL sum_loss (:646) or XE loss without weight decay (:159-166).
__global__ void loss_reduce_kernel(int mode, const float* __restrict__ logp, const float* __restrict__ sumlsm, const float* __restrict__ mask,
                                   const float* __restrict__ rewards, const float* __restrict__ base, int N, int Tc, const float* __restrict__ norm_p,
                                   float ls, int V, float* __restrict__ out, float* __restrict__ logp_masked,
                                   const float* __restrict__ colsum_global = nullptr, int n_global = 0) {
    __shared__ float red[32];
    const float norm = norm_p[0];
    float acc = 0.f;
    if (mode == 0) {
        for (int idx = threadIdx.x; idx < N * Tc; idx += blockDim.x) {
            int i = idx / N, n = idx % N;
            float m = mask[(size_t)n * Tc + i];
            float lp = logp[idx] * m;
            if (logp_masked) logp_masked[(size_t)n * Tc + i] = lp;
            acc += -(rewards[n] - base[n]) * lp;
        }
        acc = block_reduce(acc, [](float a, float b) { return a + b; }, red);
        if (threadIdx.x == 0) out[0] = acc / norm;
    } else {
        for (int i = threadIdx.x; i < Tc; i += blockDim.x) {
            float S = 0.f, ce = 0.f;
            for (int n = 0; n < N; ++n) {
                S += mask[(size_t)n * Tc + i];
                ce += -((1.f - ls) * logp[i * N + n] + (ls / (float)V) * sumlsm[i * N + n]);
            }
            if (colsum_global) S = colsum_global[i];          // this rank's share of mean_b(CE) * sum_b mask; the shares add up over ranks
            acc += ce / (float)(colsum_global ? n_global : N) * S;
        }
        acc = block_reduce(acc, [](float a, float b) { return a + b; }, red);
        if (threadIdx.x == 0) out[0] = acc / norm;
    }
}

__global__ void sum_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    acc = block_reduce(acc, [](float a, float b) { return a + b; }, red);
    if (threadIdx.x == 0) out[0] = acc;
}

// gather rows of a compute-dtype table: out[r, :] = table[idx[r], :]
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ table, int ld, const int* __restrict__ idx, int R, T* __restrict__ out) {
    int r = blockIdx.x;
    const T* s = table + (size_t)idx[r] * ld;
    T* d = out + (size_t)r * ld;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) d[c] = s[c];
}

// Embedding gradient: grad_Wemb[tok[r], e] += demb[r, e] (atomic; e < E) and the un-deduplicated slice square norm (R6).
__global__ void scatter_emb_grad_kernel(const float* __restrict__ demb, int ld, const int* __restrict__ tok, int R, int E, float* __restrict__ gW,
                                        float* __restrict__ slice_sq) {
    __shared__ float red[32];
    int r = blockIdx.x;
    float sq = 0.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float v = demb[(size_t)r * ld + e];
        sq += v * v;
        atomicAdd(gW + (size_t)tok[r] * E + e, v);
    }
    sq = block_reduce(sq, [](float a, float b) { return a + b; }, red);
    if (threadIdx.x == 0) atomicAdd(slice_sq, sq);
}

// Bias gradients: row sums of a transposed gradient matrix XT [C, ld] over the first R columns.
// gate_h > 0: row c is packed gate order 4u+g -> grad[g*gate_h + u]; else grad[c], c < ncols.
__device__ __forceinline__ float sum8(const bf16* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
    return (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
}
__device__ __forceinline__ float sum8(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    return (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
}
template <typename T>
__global__ void rowsum_grad_kernel(const T* __restrict__ XT, int ld, int R, int ncols, int gate_h, float* __restrict__ grad) {
    __shared__ float red[32];
    int c = blockIdx.x;
    int dst;
    if (gate_h > 0) { int u = c >> 2, g = c & 3; if (u >= gate_h) return; dst = g * gate_h + u; }
    else { if (c >= ncols) return; dst = c; }
    const T* row = XT + (size_t)c * ld;      // ld is a multiple of 128 elements and the pad columns are zero: vector loads over ru(R, 8)
    float acc = 0.f;
    const int R8 = (R + 7) & ~7;
    for (int r = threadIdx.x * 8; r < R8; r += blockDim.x * 8) acc += sum8(row + r);
    acc = block_reduce(acc, [](float a, float b) { return a + b; }, red);
    if (threadIdx.x == 0) grad[dst] += acc;
}

// Bias gradients straight from a row-major gradient matrix Y [R, ld]: column sums over the R rows (coalesced 2-column
// loads; rows split over blockIdx.y, partial sums combined with atomicAdd).  Column mapping as in rowsum_grad_kernel.
template <typename T>
__global__ void colsum_grad_kernel(const T* __restrict__ Y, int ld, int R, int ncols, int gate_h, float* __restrict__ grad) {
    __shared__ float red[8][64];
    const int c = blockIdx.x * 64 + threadIdx.x * 2;
    const int rows_per = (R + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
    float a0 = 0.f, a1 = 0.f;
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {       // four independent rows in flight per thread (one at a time left the kernel at a third of the HBM bandwidth)
        const T* q0 = Y + (size_t)r * ld + c; const T* q1 = q0 + (size_t)8 * ld; const T* q2 = q1 + (size_t)8 * ld; const T* q3 = q2 + (size_t)8 * ld;
        const float x0 = to_f32(q0[0]), y0 = to_f32(q0[1]), x1 = to_f32(q1[0]), y1 = to_f32(q1[1]);
        const float x2 = to_f32(q2[0]), y2 = to_f32(q2[1]), x3 = to_f32(q3[0]), y3 = to_f32(q3[1]);
        a0 += (x0 + x1) + (x2 + x3); a1 += (y0 + y1) + (y2 + y3);
    }
    for (; r < r1; r += 8) {
        const T* q = Y + (size_t)r * ld + c;
        a0 += to_f32(q[0]); a1 += to_f32(q[1]);
    }
    red[threadIdx.y][threadIdx.x * 2] = a0; red[threadIdx.y][threadIdx.x * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float acc = 0.f;
            for (int j = 0; j < 8; ++j) acc += red[j][threadIdx.x * 2 + e];
            int cc = c + e, dst;
            if (gate_h > 0) { int u = cc >> 2, g = cc & 3; if (u >= gate_h) continue; dst = g * gate_h + u; }
            else { if (cc >= ncols) continue; dst = cc; }
            atomicAdd(grad + dst, acc);
        }
    }
}

// ---- optimiser --------------------------------------------------------------------------------------------------
// sum of squares of a float range into a double accumulator (global-norm clip)
// 128-bit loads, four in flight per thread: a scalar grid-stride loop keeps too few bytes in flight to reach HBM bandwidth.
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    size_t head = ((16 - ((uintptr_t)g & 15)) & 15) / 4;
    if (head > n) head = n;
    const float4* g4 = reinterpret_cast<const float4*>(g + head);
    const size_t n4 = (n - head) / 4, stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 a = g4[i], b = g4[i + stride], c = g4[i + 2 * stride], d = g4[i + 3 * stride];
        acc += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
        acc += (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z + (double)b.w * b.w;
        acc += (double)c.x * c.x + (double)c.y * c.y + (double)c.z * c.z + (double)c.w * c.w;
        acc += (double)d.x * d.x + (double)d.y * d.y + (double)d.z * d.z + (double)d.w * d.w;
    }
    for (; i < n4; i += stride) { const float4 a = g4[i]; acc += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w; }
    if (blockIdx.x == 0) {   // unaligned head and the < 4 element tail
        for (size_t j = threadIdx.x; j < head; j += blockDim.x) { float v = g[j]; acc += (double)v * v; }
        for (size_t j = head + 4 * n4 + threadIdx.x; j < n; j += blockDim.x) { float v = g[j]; acc += (double)v * v; }
    }
    acc = block_reduce(acc, [](double a, double b) { return a + b; }, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// TF-1.1 Adam (R7) with the global-norm clip scale folded in (tf.clip_by_global_norm [lib]):
//   scale = clip * min(1/gn, 1/clip);  g' = g*scale + l2*theta;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
//   theta -= lr_t * m / (sqrt(v) + eps),  lr_t = lr * sqrt(1-b2^t)/(1-b1^t) (computed on the host)
// sq[0] = dense sum of squares of all gradients, sq[1] = dense Wemb part, sq[2] = Wemb slice square norm.
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                            const double* __restrict__ sq, int use_slice_norm, int normalize, float clip, float lr_t, float b1, float b2, float eps,
                            float* __restrict__ gnorm_out) {
    // normalize: the gradients were accumulated with norm = 1 (data-parallel); g[n+2] holds the global sum(mask) (R1)
    const float inv = normalize ? 1.0f / g[n + 2] : 1.0f;
    double s = sq[0];
    if (use_slice_norm) s = s - sq[1] + sq[2];
    s *= (double)inv * (double)inv;
    float gn = (float)sqrt(s);
    float scale = 1.0f;
    if (clip > 0.f && gn > 0.f) scale = clip * fminf(1.0f / gn, 1.0f / clip);
    if (gnorm_out && blockIdx.x == 0 && threadIdx.x == 0) { gnorm_out[0] = gn; gnorm_out[1] = g[n + 1] * inv; }
    // the four blocks are 256-byte aligned arena allocations: 128-bit accesses over the bulk, scalars over the < 4 element tail
    const size_t n4 = n / 4, stride = (size_t)gridDim.x * blockDim.x;
    float4* t4 = reinterpret_cast<float4*>(theta); const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 gg = g4[i]; float4 mm = m4[i], vv = v4[i], tt = t4[i];
        float gi;
        gi = gg.x * inv * scale; mm.x = b1 * mm.x + (1.f - b1) * gi; vv.x = b2 * vv.x + (1.f - b2) * gi * gi; tt.x -= lr_t * mm.x / (sqrtf(vv.x) + eps);
        gi = gg.y * inv * scale; mm.y = b1 * mm.y + (1.f - b1) * gi; vv.y = b2 * vv.y + (1.f - b2) * gi * gi; tt.y -= lr_t * mm.y / (sqrtf(vv.y) + eps);
        gi = gg.z * inv * scale; mm.z = b1 * mm.z + (1.f - b1) * gi; vv.z = b2 * vv.z + (1.f - b2) * gi * gi; tt.z -= lr_t * mm.z / (sqrtf(vv.z) + eps);
        gi = gg.w * inv * scale; mm.w = b1 * mm.w + (1.f - b1) * gi; vv.w = b2 * vv.w + (1.f - b2) * gi * gi; tt.w -= lr_t * mm.w / (sqrtf(vv.w) + eps);
        m4[i] = mm; v4[i] = vv; t4[i] = tt;
    }
    for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * inv * scale;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// L2 weight decay gradient (Q4): g += decay * theta over a range
__global__ void add_decay_kernel(float* __restrict__ g, const float* __restrict__ theta, size_t n, float decay) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) g[i] += decay * theta[i];
}
