// Temporal-attention caption decoder (BASELINE config 3): the model of original_attention.py:54-251 --
// frame projection, additive attention over the frame embeddings, single LSTM3, tanh MLP head, vocabulary projection --
// as the greedy sampler (build_generator :155-199 / build_sampler :201-251, with the saved alphas), the teacher-forced
// loss of build_model (:88-152, DropoutWrapper on the LSTM3 output, hinge regulariser on the first 8 alphas), its
// gradients (back-propagation through the 35 decode steps) and the optimiser of train() (:427-435: global-norm clip 10,
// TF Adam).
// Backward structure: everything that does not feed the recurrence is batched over time -- d logits / dWo / d head / dWp
// before the reverse loop, the weight gradients of LSTM3, Wa, Ua, We after it (single GEMMs over the stashed per-step
// operands); the loop itself carries only the cell backward, dg . W3^T, the attention backward and dq . Wa^T.
//
// Contractions run on the library's GEMM mainloops (tcgen05 / TMA / TMEM for bf16, SIMT FMA for the fp32 mode) with the
// EpiStore epilogue; the per-step glue (attention scores + softmax + weighted frame sum, LSTM cell, head non-linearity,
// arg-max, embedding gather, cross entropy) is the small kernels below.  Per decode step and row the attention reads the
// video's [n, H] `image_part` and `image_emb` tables (fp32) once: 2 n H 4 B, HBM/L2-bound.
// Layout: frames are video-major rows b * n + i (the reference's [n, b, h] transpose is only a view); LSTM3 gate columns
// keep the TF order g * Hp + u; the concatenated operands [atten | emb | h] and [out1 | atten | emb] are kept as
// three Hp-wide blocks of one row so each product is ONE GEMM with K = 3 Hp.
#include <algorithm>
#include <type_traits>

#include "engine.cuh"
#include "gemm.cuh"
#include "gemm_tcgen05.cuh"

struct AttVar { std::string name; std::vector<std::string> aliases; int64_t rows, cols; size_t off; size_t count() const { return (size_t)rows * (cols ? cols : 1); } };

struct s2vt_att_handle {
    s2vt_att_config cfg;
    int D, H, V, n, Tc, Dp, Hp, Vp;
    size_t esz, P;
    std::vector<AttVar> vars;
    char* state = nullptr; size_t state_bytes = 0;
    char* ws = nullptr; size_t ws_bytes = 0;
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;   // grads: P floats + 4 aux (slice sq-norm, loss, reg, sum mask)
    double* sq = nullptr;
    void *WeT, *UaT, *WaT, *W3T, *WpT, *WoT;      // forward operands [N, K] (K-major)
    void *UaN, *WaN, *W3N, *WpN, *WoN;            // backward operands: the TF layouts, block-padded ([K_fwd, N_fwd] = [N_bwd, K_bwd])
    bool train_valid = false; int train_B = 0;
    float *be_p, *ba_p, *w_p, *b3_p, *bp_p, *bo_p;
    bool bound = false, fresh = false;
    long long launches = 0;
    tc::MapCache* maps = nullptr;
    int iWemb, iWe, ibe, iw, iWa, iUa, iba, iWo, ibo, iWp, ibp, iW3, ib3;
    mutable std::string err;
    int fail(int code, const char* fmt, ...) const {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
    float* P_(int i) const { return params + vars[i].off; }
};

#define ATRY(x) do { int _r = (x); if (_r) return _r; } while (0)

// ---- kernels ---------------------------------------------------------------------------------------------------------
// dst[(cb * Cbp + c) * ldd + rb * Rbp + r] = src[(rb * Rb + r) * lds + cb * Cb + c]: TF-layout matrix whose rows / columns are
// blocks of Rb / Cb (concatenated inputs, gate blocks) -> K-major operand with every block padded to Rbp / Cbp.
// Per-step element-wise kernels let the GEMM that follows (a programmatic dependent, att_gemm) start its prologue at once; that
// GEMM touches activations only after griddepcontrol.wait, i.e. after this kernel has completed.
#define ATT_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")

template <typename T>
__global__ void att_pack_kernel(const float* __restrict__ src, int lds, int R, int C, T* __restrict__ dst, int ldd, int Rb, int Cb, int Rbp, int Cbp, int transpose) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * C; idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / C), c = (int)(idx % C);
        const int rp = (r / Rb) * Rbp + r % Rb, cp = (c / Cb) * Cbp + c % Cb;
        dst[transpose ? (size_t)cp * ldd + rp : (size_t)rp * ldd + cp] = from_f32<T>(src[(size_t)r * lds + c]);
    }
}
// inverse for gradients: grad[r * ldg + c] += padded[rp * ldp + cp]
__global__ void att_unpack_add_kernel(const float* __restrict__ padded, int ldp, int R, int C, float* __restrict__ grad, int Rb, int Cb, int Rbp, int Cbp) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * C; idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / C), c = (int)(idx % C);
        grad[idx] += padded[(size_t)((r / Rb) * Rbp + r % Rb) * ldp + (c / Cb) * Cbp + c % Cb];
    }
}
__global__ void att_pack_vec_kernel(const float* __restrict__ src, int C, float* __restrict__ dst, int Cb, int Cbp) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) dst[(c / Cb) * Cbp + c % Cb] = src[c];
}
template <typename T>
__global__ void att_convert_video_kernel(const float* __restrict__ video, size_t rows, int D, int Dp, T* __restrict__ out) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < rows * Dp; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / Dp; int d = (int)(idx % Dp);
        out[idx] = from_f32<T>(d < D ? video[r * D + d] : 0.f);
    }
}

// :113-128 for one row: e_i = sum_h tanh(q_h + part[v, i, h]) w_h ; alpha = exp(e) / (sum exp(e) (+1 if 0)) ; atten_h = sum_i alpha_i emb[v, i, h].
// Writes atten into the LSTM3 operand (block 0 of x3) and the head operand (block 1 of x4), the alphas ([T_c, n, B] layout of
// saved_alphas, :237, nullable) and the hinge term max(0, m - sum_{i < 8} alpha_i) of the regulariser (:146).
template <typename T>
__global__ void __launch_bounds__(256) att_attend_kernel(const float* __restrict__ q, const float* __restrict__ part, const float* __restrict__ emb,
                                                         const float* __restrict__ w, int B, int n, int Hp, T* __restrict__ x3, T* __restrict__ x4,
                                                         float* __restrict__ alphas_out, int R, float m_hinge, int reg_frames, float* __restrict__ hinge, float* __restrict__ alpha_rows) {
    ATT_PDL_TRIGGER();
    __shared__ float e[128];
    __shared__ float red[32];
    const int r = blockIdx.x, v = r % B, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const float* qr = q + (size_t)r * Hp;
    for (int i = warp; i < n; i += nw) {
        const float* pr = part + ((size_t)v * n + i) * Hp;
        float s = 0.f;
        for (int h = lane * 4; h < Hp; h += 128) {
            float4 a = *reinterpret_cast<const float4*>(qr + h), b = *reinterpret_cast<const float4*>(pr + h), ww = *reinterpret_cast<const float4*>(w + h);
            s += tanh_<T>(a.x + b.x) * ww.x + tanh_<T>(a.y + b.y) * ww.y + tanh_<T>(a.z + b.z) * ww.z + tanh_<T>(a.w + b.w) * ww.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) e[i] = expf(s);
    }
    __syncthreads();
    float part_sum = 0.f;
    for (int i = tid; i < n; i += blockDim.x) part_sum += e[i];
    float denom = block_reduce(part_sum, [](float a, float b) { return a + b; }, red);
    if (denom == 0.f) denom = 1.f;
    const float inv = 1.0f / denom;
    if (tid == 0) {
        float s8 = 0.f;
        for (int i = 0; i < n && i < reg_frames; ++i) s8 += e[i] * inv;
        if (hinge) hinge[r] = fmaxf(0.f, m_hinge - s8);
    }
    if (alphas_out)
        for (int i = tid; i < n; i += blockDim.x) alphas_out[(size_t)i * R + r] = e[i] * inv;
    if (alpha_rows)
        for (int i = tid; i < n; i += blockDim.x) alpha_rows[(size_t)r * n + i] = e[i] * inv;
    for (int h = tid; h < Hp; h += blockDim.x) {
        float a = 0.f;
        for (int i = 0; i < n; ++i) a += e[i] * inv * emb[((size_t)v * n + i) * Hp + h];
        T t = from_f32<T>(a);
        x3[(size_t)r * 3 * Hp + h] = t;
        x4[(size_t)r * 3 * Hp + Hp + h] = t;
    }
}

// BasicLSTMCell on the pre-activations g [R, 4Hp] (TF gate order i, j, f, o in blocks of Hp, bias already added): new state into
// c_out, un-dropped h into block 2 of the NEXT step's x3 (the cell's own recurrent input), DropoutWrapper output into the next
// step's query operand hq (h_prev = output1 :135) and block 0 of this step's x4 (head operand).
template <typename T>
__global__ void att_cell_kernel(const float* __restrict__ g, const float* __restrict__ c_in, float* __restrict__ c_out, int R, int Hp, T* __restrict__ x3_next,
                                T* __restrict__ x4, T* __restrict__ hq_next, unsigned long long seed, uint32_t step, uint32_t row_base, float keep) {
    ATT_PDL_TRIGGER();
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * Hp; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / Hp), u = (int)(idx % Hp);
        const float* gr = g + (size_t)r * 4 * Hp;
        const float si = sigm<T>(gr[u]), tj = tanh_<T>(gr[Hp + u]), sf = sigm<T>(gr[2 * Hp + u] + 1.0f), so = sigm<T>(gr[3 * Hp + u]);
        const float cn = c_in[idx] * sf + si * tj;
        const float hn = tanh_<T>(cn) * so;
        c_out[idx] = cn;
        x3_next[(size_t)r * 3 * Hp + 2 * Hp + u] = from_f32<T>(hn);
        const float out = keep < 1.0f ? hn * dropout_mult(seed, S2VT_STREAM_DROP1, row_base + (uint32_t)r, step, (uint32_t)u, keep) : hn;
        hq_next[idx] = from_f32<T>(out);
        x4[(size_t)r * 3 * Hp + u] = from_f32<T>(out);
    }
}
template <typename T>
__global__ void att_tanh_kernel(const float* __restrict__ x, size_t count, T* __restrict__ out, float* __restrict__ outF) {
    ATT_PDL_TRIGGER();
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) {
        const float t = tanh_<T>(x[idx]);
        out[idx] = from_f32<T>(t);
        if (outF) outF[idx] = t;
    }
}
// tf.argmax(logit_words, 1) (:190): lowest index among equal maxima; also the id matrix column.
__global__ void __launch_bounds__(256) att_argmax_kernel(const float* __restrict__ logits, int ld, int V, int* __restrict__ tok, int* __restrict__ ids, int Tc, int t) {
    ATT_PDL_TRIGGER();
    __shared__ ArgVal red[32];
    const int r = blockIdx.x;
    ArgVal best; best.v = -INFINITY; best.i = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += blockDim.x) { ArgVal c; c.v = logits[(size_t)r * ld + v]; c.i = v; best = argmax_op(best, c); }
    best = block_argmax(best, red);
    if (threadIdx.x == 0) { tok[r] = best.i; ids[(size_t)r * Tc + t] = best.i; }
}
// current_embed = Wemb[word] (:148, 192) into block 1 of the next x3 and block 2 of the next x4; word = tok[r] or caption[r, t].
template <typename T>
__global__ void att_gather_kernel(const float* __restrict__ Wemb, int H, int Hp, const int* __restrict__ tok, int tok_ld, int tok_col, int R, T* __restrict__ x3,
                                  T* __restrict__ x4) {
    ATT_PDL_TRIGGER();
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * H; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / H), e = (int)(idx % H);
        const T t = from_f32<T>(Wemb[(size_t)tok[(size_t)r * tok_ld + tok_col] * H + e]);
        x3[(size_t)r * 3 * Hp + Hp + e] = t;
        x4[(size_t)r * 3 * Hp + 2 * Hp + e] = t;
    }
}
// softmax_cross_entropy_with_logits against the one-hot label, times the mask, plus beta * hinge * mask (:144-148).
// acc[0] += sum(ce * mask + reg), acc[1] += sum(reg), acc[2] += sum(mask)
__global__ void __launch_bounds__(256) att_ce_kernel(const float* __restrict__ logits, int ld, int V, const int* __restrict__ cap, const float* __restrict__ mask, int Tc,
                                                     int t, const float* __restrict__ hinge, float beta, double* __restrict__ acc) {
    ATT_PDL_TRIGGER();
    __shared__ float red[32];
    const int r = blockIdx.x;
    const float* row = logits + (size_t)r * ld;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, row[v]);
    mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(row[v] - mx);
    s = block_reduce(s, [](float a, float b) { return a + b; }, red);
    if (threadIdx.x == 0) {
        const float mk = mask[(size_t)r * Tc + t];
        const float ce = (mx + logf(s)) - row[cap[(size_t)r * Tc + t]];
        const float reg = beta * hinge[r] * mk;
        atomicAdd(&acc[0], (double)(ce * mk + reg));
        atomicAdd(&acc[1], (double)reg);
        atomicAdd(&acc[2], (double)mk);
    }
}
__global__ void att_loss_final_kernel(const double* __restrict__ acc, float* __restrict__ out, float* __restrict__ aux) {
    const float loss = (float)(acc[0] / acc[2]), reg = (float)(acc[1] / acc[2]);
    if (out) { out[0] = loss; out[1] = reg; }
    if (aux) { aux[1] = loss; aux[2] = reg; aux[3] = (float)acc[2]; }
}

// ---- backward kernels ------------------------------------------------------------------------------------------------
// d loss / d logits of every step at once: (softmax(z) - onehot(y)) * mask / sum(mask), padded columns zero.  Row = t * R + r.
template <typename T>
__global__ void __launch_bounds__(256) att_dlogits_kernel(const float* __restrict__ logits, int Vp, int V, const int* __restrict__ cap, const float* __restrict__ mask,
                                                          int R, int Tc, const double* __restrict__ acc, T* __restrict__ dz) {
    __shared__ float red[32];
    const int row = blockIdx.x, t = row / R, r = row % R;
    const float* z = logits + (size_t)row * Vp;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, z[v]);
    mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(z[v] - mx);
    s = block_reduce(s, [](float a, float b) { return a + b; }, red);
    const float coef = mask[(size_t)r * Tc + t] / (float)acc[2], inv = 1.0f / s;
    const int y = cap[(size_t)r * Tc + t];
    for (int v = threadIdx.x; v < Vp; v += blockDim.x)
        dz[(size_t)row * Vp + v] = from_f32<T>(v < V ? (expf(z[v] - mx) * inv - (v == y ? 1.0f : 0.0f)) * coef : 0.0f);
}
// d head pre-activation: du = do2 * (1 - o2^2)
template <typename T>
__global__ void att_du_kernel(const float* __restrict__ do2, const float* __restrict__ o2, size_t count, T* __restrict__ du) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) {
        const float o = o2[idx];
        du[idx] = from_f32<T>(do2[idx] * (1.0f - o * o));
    }
}
// LSTM3 cell backward at step t.  do1 = dx4[:, block 0] (+ dhq_next: the next step's query reads output1); dh = do1 * dropout
// multiplier + dh_rec (block 2 of the next step's dx3); gates are recomputed from the stashed pre-activations.
template <typename T>
__global__ void att_cell_bwd_kernel(const float* __restrict__ dx4, const float* __restrict__ dhq_next, const float* __restrict__ dx3_next, const float* __restrict__ g,
                                    const float* __restrict__ c_prev, const float* __restrict__ c_new, float* __restrict__ dc, int R, int Hp, T* __restrict__ dg,
                                    unsigned long long seed, uint32_t step, uint32_t row_base, float keep) {
    ATT_PDL_TRIGGER();
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * Hp; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / Hp), u = (int)(idx % Hp);
        float do1 = dx4[(size_t)r * 3 * Hp + u];
        if (dhq_next) do1 += dhq_next[idx];
        if (keep < 1.0f) do1 *= dropout_mult(seed, S2VT_STREAM_DROP1, row_base + (uint32_t)r, step, (uint32_t)u, keep);
        const float dh = do1 + (dx3_next ? dx3_next[(size_t)r * 3 * Hp + 2 * Hp + u] : 0.0f);
        const float* gr = g + (size_t)r * 4 * Hp;
        const float si = sigm<T>(gr[u]), tj = tanh_<T>(gr[Hp + u]), sf = sigm<T>(gr[2 * Hp + u] + 1.0f), so = sigm<T>(gr[3 * Hp + u]);
        const float tc = tanh_<T>(c_new[idx]);
        const float d_o = dh * tc;
        const float dcn = dc[idx] + dh * so * (1.0f - tc * tc);
        dc[idx] = dcn * sf;
        T* d = dg + (size_t)r * 4 * Hp;
        d[u] = from_f32<T>(dcn * tj * si * (1.0f - si));
        d[Hp + u] = from_f32<T>(dcn * si * (1.0f - tj * tj));
        d[2 * Hp + u] = from_f32<T>(dcn * c_prev[idx] * sf * (1.0f - sf));
        d[3 * Hp + u] = from_f32<T>(d_o * so * (1.0f - so));
    }
}
// Attention backward for one row (one CTA per row; rows of a call belong to distinct videos, so the per-video accumulators need
// no atomics).  da = dx4[:, block 1] + dx3[:, block 0]; d alpha_i = <da, emb_i> (- beta mask / norm on the first 8 frames when the
// hinge is active); de = alpha (d alpha - <alpha, d alpha>); d pre_ih = de_i w_h (1 - s_ih^2) with s = tanh(q + part) recomputed;
// dq_h = sum_i d pre_ih; d part += d pre; d emb_i += alpha_i da; dw_h += sum_i de_i s_ih.
template <typename T>
__global__ void __launch_bounds__(256) att_attend_bwd_kernel(const float* __restrict__ dx4, const float* __restrict__ dx3, const float* __restrict__ q,
                                                             const float* __restrict__ part, const float* __restrict__ emb, const float* __restrict__ w,
                                                             const float* __restrict__ alpha_rows, const float* __restrict__ hinge, const float* __restrict__ mask,
                                                             int Tc, int t, const double* __restrict__ acc, float beta, int reg_frames, int B, int n, int Hp,
                                                             T* __restrict__ dq, float* __restrict__ d_part, float* __restrict__ d_emb, float* __restrict__ dw) {
    ATT_PDL_TRIGGER();
    extern __shared__ float sm[];
    float* da = sm;              // [Hp]
    float* dal = sm + Hp;        // [n] d alpha, then de
    __shared__ float red[32];
    const int r = blockIdx.x, v = r % B, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int h = tid; h < Hp; h += blockDim.x) da[h] = dx4[(size_t)r * 3 * Hp + Hp + h] + dx3[(size_t)r * 3 * Hp + h];
    __syncthreads();
    const float* al = alpha_rows + (size_t)r * n;
    const float creg = hinge[r] > 0.0f ? beta * mask[(size_t)r * Tc + t] / (float)acc[2] : 0.0f;
    for (int i = warp; i < n; i += nw) {
        const float* er = emb + ((size_t)v * n + i) * Hp;
        float s = 0.f;
        for (int h = lane; h < Hp; h += 32) s += da[h] * er[h];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) dal[i] = s - (i < reg_frames ? creg : 0.0f);
    }
    __syncthreads();
    float p = 0.f;
    for (int i = tid; i < n; i += blockDim.x) p += al[i] * dal[i];
    const float dot = block_reduce(p, [](float a, float b) { return a + b; }, red);
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) dal[i] = al[i] * (dal[i] - dot);
    __syncthreads();
    for (int h = tid; h < Hp; h += blockDim.x) {
        const float qh = q[(size_t)r * Hp + h], wh = w[h], dah = da[h];
        float dqa = 0.f, dwa = 0.f;
        for (int i = 0; i < n; ++i) {
            const size_t o = ((size_t)v * n + i) * Hp + h;
            const float s = tanh_<T>(qh + part[o]);
            const float dpre = dal[i] * wh * (1.0f - s * s);
            dqa += dpre;
            dwa += dal[i] * s;
            d_part[o] += dpre;
            d_emb[o] += al[i] * dah;
        }
        dq[(size_t)r * Hp + h] = from_f32<T>(dqa);
        atomicAdd(&dw[h], dwa);
    }
}
// Embedding gradient of step t: d current_embed = dx4[:, block 2] + dx3[:, block 1] scattered onto row caption[r, t-1] of dWemb, plus
// the un-deduplicated IndexedSlices square norm tf.clip_by_global_norm sees in sq[2].
__global__ void att_scatter_emb_kernel(const float* __restrict__ dx4, const float* __restrict__ dx3, const int* __restrict__ cap, int Tc, int t, int R, int H, int Hp,
                                       float* __restrict__ gW, double* __restrict__ sq) {
    ATT_PDL_TRIGGER();
    __shared__ double red[32];
    double acc = 0.0;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * H; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / H), e = (int)(idx % H);
        const float d = dx4[(size_t)r * 3 * Hp + 2 * Hp + e] + dx3[(size_t)r * 3 * Hp + Hp + e];
        atomicAdd(&gW[(size_t)cap[(size_t)r * Tc + t - 1] * H + e], d);
        acc += (double)d * d;
    }
    acc = block_reduce(acc, [](double a, double b) { return a + b; }, red);
    if (threadIdx.x == 0) atomicAdd(&sq[2], acc);
}
// grad[c] += sum_rows Y[row, pad(c)] (bias gradients; block-padded columns)
template <typename U>
__global__ void att_colsum_kernel(const U* __restrict__ Y, int ld, int rows, int C, int Cb, int Cbp, float* __restrict__ grad) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int cp = (c / Cb) * Cbp + c % Cb;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += to_f32(Y[(size_t)r * ld + cp]);
    grad[c] += s;
}
template <typename T>
__global__ void att_cast_kernel(const float* __restrict__ x, size_t count, T* __restrict__ out) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) out[idx] = from_f32<T>(x[idx]);
}
// dst[c * ldd + r] = src[r * lds + c] (fp32-mode weight gradients need K-major operands; dst is zero-filled beforehand)
template <typename T>
__global__ void att_transpose_kernel(const T* __restrict__ src, int lds, int R, int C, T* __restrict__ dst, int ldd) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * C; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / C), c = (int)(idx % C);
        dst[(size_t)c * ldd + r] = src[(size_t)r * lds + c];
    }
}
__global__ void att_sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float v = g[i]; acc += (double)v * v; }
    acc = block_reduce(acc, [](double a, double b) { return a + b; }, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}
// tf.clip_by_global_norm(gradients, clip) + tf.train.AdamOptimizer apply (:431-435).  sq[0] dense sum of squares of all gradients,
// sq[1] dense Wemb part, sq[2] Wemb IndexedSlices square norm (what the reference's global norm uses for Wemb).
__global__ void att_adam_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                                const double* __restrict__ sq, float clip, float lr_t, float b1, float b2, float eps, float* __restrict__ out) {
    const float gn = (float)sqrt(sq[0] - sq[1] + sq[2]);
    float scale = 1.0f;
    if (clip > 0.f && gn > 0.f) scale = clip * fminf(1.0f / gn, 1.0f / clip);
    if (out && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = gn; out[1] = g[n + 1]; }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// ---- host ------------------------------------------------------------------------------------------------------------
static void att_add_var(s2vt_att_handle* h, const char* name, std::vector<std::string> aliases, int64_t rows, int64_t cols) {
    AttVar v; v.name = name; v.aliases = aliases; v.rows = rows; v.cols = cols; v.off = h->P;
    h->P += v.count();
    h->vars.push_back(v);
}

template <typename F> static void att_layout(s2vt_att_handle* h, Arena& a, F assign) {
    const size_t e = h->esz;
    const int Dp = h->Dp, Hp = h->Hp, Vp = h->Vp;
    assign(0, a.take<float>(h->P));
    assign(13, a.take<float>(h->P + 8)); assign(14, a.take<float>(h->P)); assign(15, a.take<float>(h->P)); assign(16, a.take<double>(4));
    assign(1, a.take<char>((size_t)Hp * Dp * e)); assign(2, a.take<char>((size_t)Hp * Hp * e)); assign(3, a.take<char>((size_t)Hp * Hp * e));
    assign(4, a.take<char>((size_t)4 * Hp * 3 * Hp * e)); assign(5, a.take<char>((size_t)Hp * 3 * Hp * e)); assign(6, a.take<char>((size_t)Vp * Hp * e));
    assign(17, a.take<char>((size_t)Hp * Hp * e)); assign(18, a.take<char>((size_t)Hp * Hp * e)); assign(19, a.take<char>((size_t)3 * Hp * 4 * Hp * e));
    assign(20, a.take<char>((size_t)3 * Hp * Hp * e)); assign(21, a.take<char>((size_t)Hp * Vp * e));
    assign(7, a.take<float>(Hp)); assign(8, a.take<float>(Hp)); assign(9, a.take<float>(Hp)); assign(10, a.take<float>(4 * Hp)); assign(11, a.take<float>(Hp));
    assign(12, a.take<float>(Vp));
}

extern "C" int s2vt_att_create(const s2vt_att_config* cfg, s2vt_att_handle** out) {
    if (!cfg || !out) return S2VT_EINVAL;
    if (cfg->dim_image <= 0 || cfg->dim_hidden <= 0 || cfg->n_words <= 2 || cfg->n_words > 65000 || cfg->n_video_steps <= 0 || cfg->n_video_steps > 128 ||
        cfg->n_caption_steps <= 0 || (cfg->precision != S2VT_PREC_BF16 && cfg->precision != S2VT_PREC_FP32) || !(cfg->dropout_keep > 0.f && cfg->dropout_keep <= 1.f))
        return S2VT_EINVAL;
    s2vt_att_handle* h = new s2vt_att_handle();
    h->cfg = *cfg;
    h->D = cfg->dim_image; h->H = cfg->dim_hidden; h->V = cfg->n_words; h->n = cfg->n_video_steps; h->Tc = cfg->n_caption_steps;
    h->Dp = ru(h->D, S2VT_PAD); h->Hp = ru(h->H, S2VT_PAD); h->Vp = ru(h->V, 256);
    h->esz = cfg->precision == S2VT_PREC_BF16 ? 2 : 4;
    h->P = 0;
    const int H = h->H;
    // variables of original_attention.py:64-86 by their TF names
    h->iWemb = 0; att_add_var(h, "Wemb", {}, h->V, H);
    h->iWe = 1; att_add_var(h, "encode_image_W", {}, h->D, H);
    h->ibe = 2; att_add_var(h, "encode_image_b", {}, H, 0);
    h->iw = 3; att_add_var(h, "embed_att_w", {}, H, 1);
    h->iWa = 4; att_add_var(h, "embed_att_Wa", {}, H, H);
    h->iUa = 5; att_add_var(h, "embed_att_Ua", {}, H, H);
    h->iba = 6; att_add_var(h, "embed_att_ba", {}, H, 0);
    h->iWo = 7; att_add_var(h, "embed_word_W", {}, H, h->V);
    h->ibo = 8; att_add_var(h, "embed_word_b", {}, h->V, 0);
    h->iWp = 9; att_add_var(h, "embed_nn_Wp", {}, 3 * H, H);
    h->ibp = 10; att_add_var(h, "embed_nn_bp", {}, H, 0);
    h->iW3 = 11; att_add_var(h, "s2vt/LSTM3/basic_lstm_cell/weights", {"s2vt/LSTM3/basic_lstm_cell/kernel", "s2vt/LSTM3/BasicLSTMCell/Linear/Matrix"}, 3 * H, 4 * H);
    h->ib3 = 12; att_add_var(h, "s2vt/LSTM3/basic_lstm_cell/biases", {"s2vt/LSTM3/basic_lstm_cell/bias", "s2vt/LSTM3/BasicLSTMCell/Linear/Bias"}, 4 * H, 0);
    Arena a(nullptr, 0);
    att_layout(h, a, [](int, void*) {});
    h->state_bytes = a.used;
    *out = h;
    return S2VT_OK;
}
extern "C" void s2vt_att_destroy(s2vt_att_handle* h) { if (h) { delete h->maps; delete h; } }
extern "C" const char* s2vt_att_last_error(const s2vt_att_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" size_t s2vt_att_num_params(const s2vt_att_handle* h) { return h->P; }
extern "C" size_t s2vt_att_state_bytes(const s2vt_att_handle* h) { return h->state_bytes; }
extern "C" float* s2vt_att_params(const s2vt_att_handle* h) { return h->params; }
extern "C" float* s2vt_att_grads(const s2vt_att_handle* h) { return h->grads; }
extern "C" float* s2vt_att_adam_m(const s2vt_att_handle* h) { return h->adam_m; }
extern "C" float* s2vt_att_adam_v(const s2vt_att_handle* h) { return h->adam_v; }
extern "C" int s2vt_att_num_variables(const s2vt_att_handle* h) { return (int)h->vars.size(); }
extern "C" int s2vt_att_variable_info(const s2vt_att_handle* h, int index, const char** tf_name, int64_t* offset, int64_t shape[2], int* ndim) {
    if (!h || index < 0 || index >= (int)h->vars.size()) return S2VT_EINVAL;
    const AttVar& v = h->vars[index];
    if (tf_name) *tf_name = v.name.c_str();
    if (offset) *offset = (int64_t)v.off;
    if (shape) { shape[0] = v.rows; shape[1] = v.cols; }
    if (ndim) *ndim = v.cols ? 2 : 1;
    return S2VT_OK;
}

// per-call scratch for B videos (one row per video).  train: the per-step operands are stashed for the backward pass.
struct AttWork {
    void *videoT, *embT, *hq, *x3, *x4, *oT;           // forward (single-step buffers; with train they are [T_c] arrays: x3, x4, hq, oT)
    float *emb, *part, *q, *g, *c, *o, *logits, *hinge, *alph;
    int* tok; double* acc;
    // train only
    float *o2F, *do2, *dx4, *dx3a, *dx3b, *dhq, *dc, *d_part, *d_emb, *dw, *tmpW;
    void *dz, *du, *dg, *dq, *d_partT, *d_embT, *tA, *tB;
};
static void att_plan(const s2vt_att_handle* h, Arena& a, int B, bool train, AttWork& w) {
    const size_t e = h->esz, F = (size_t)B * h->n, R = (size_t)B, S = train ? (size_t)h->Tc : 1;
    const size_t Hp = h->Hp, Vp = h->Vp, Dp = h->Dp;
    w.videoT = a.take<char>(F * Dp * e); w.embT = a.take<char>(F * Hp * e);
    w.emb = a.take<float>(F * Hp); w.part = a.take<float>(F * Hp);
    w.hq = a.take<char>((S + 1) * R * Hp * e); w.x3 = a.take<char>((S + 1) * R * 3 * Hp * e); w.x4 = a.take<char>((S + 1) * R * 3 * Hp * e);
    w.oT = a.take<char>(S * R * Hp * e);
    w.q = a.take<float>(S * R * Hp); w.g = a.take<float>(S * R * 4 * Hp); w.c = a.take<float>((S + 1) * R * Hp); w.o = a.take<float>(R * Hp);
    w.logits = a.take<float>(S * R * Vp); w.hinge = a.take<float>(S * R); w.alph = a.take<float>(S * R * h->n);
    w.tok = a.take<int>(R); w.acc = a.take<double>(4);
    if (!train) return;
    const size_t TR = S * R, TRp = ru64(TR, 16);
    w.o2F = a.take<float>(TR * Hp); w.do2 = a.take<float>(TR * Hp); w.dx4 = a.take<float>(TR * 3 * Hp);
    w.dx3a = a.take<float>(R * 3 * Hp); w.dx3b = a.take<float>(R * 3 * Hp); w.dhq = a.take<float>(R * Hp); w.dc = a.take<float>(R * Hp);
    w.d_part = a.take<float>(F * Hp); w.d_emb = a.take<float>(F * Hp); w.dw = a.take<float>(Hp);
    const size_t wmax = std::max(std::max(Hp * Vp, 3 * Hp * 4 * Hp), Dp * Hp);
    w.tmpW = a.take<float>(wmax);
    w.dz = a.take<char>(TR * Vp * e); w.du = a.take<char>(TR * Hp * e); w.dg = a.take<char>(TR * 4 * Hp * e); w.dq = a.take<char>(TR * Hp * e);
    w.d_partT = a.take<char>(F * Hp * e); w.d_embT = a.take<char>(F * Hp * e);
    if (h->cfg.precision == S2VT_PREC_FP32) {      // transposed operands of the SIMT weight-gradient GEMMs
        const size_t rows = std::max(TRp, ru64(F, 16));
        w.tA = a.take<char>(std::max(3 * Hp, Dp) * rows * e); w.tB = a.take<char>(std::max(Vp, 4 * Hp) * rows * e);
    } else { w.tA = w.tB = nullptr; }
}
extern "C" size_t s2vt_att_workspace_bytes(const s2vt_att_handle* h, int n_videos, int train) {
    if (!h || n_videos <= 0) return 0;
    Arena a(nullptr, 0);
    AttWork w;
    att_plan(h, a, n_videos, train != 0, w);
    return a.used;
}
extern "C" int s2vt_att_bind(s2vt_att_handle* h, void* state, size_t state_bytes, void* workspace, size_t workspace_bytes) {
    if (!h || !state || !workspace) return S2VT_EINVAL;
    if (state_bytes < h->state_bytes) return h->fail(S2VT_ENOSPACE, "state block holds %zu bytes, %zu needed", state_bytes, h->state_bytes);
    if (((uintptr_t)state | (uintptr_t)workspace) & 255) return h->fail(S2VT_EINVAL, "state and workspace must be 256-byte aligned");
    h->state = (char*)state; h->ws = (char*)workspace; h->ws_bytes = workspace_bytes;
    Arena a(state, state_bytes);
    att_layout(h, a, [&](int slot, void* p) {
        switch (slot) {
            case 0: h->params = (float*)p; break;
            case 1: h->WeT = p; break; case 2: h->UaT = p; break; case 3: h->WaT = p; break; case 4: h->W3T = p; break; case 5: h->WpT = p; break;
            case 6: h->WoT = p; break; case 7: h->be_p = (float*)p; break; case 8: h->ba_p = (float*)p; break; case 9: h->w_p = (float*)p; break;
            case 10: h->b3_p = (float*)p; break; case 11: h->bp_p = (float*)p; break; case 12: h->bo_p = (float*)p; break;
            case 13: h->grads = (float*)p; break; case 14: h->adam_m = (float*)p; break; case 15: h->adam_v = (float*)p; break; case 16: h->sq = (double*)p; break;
            case 17: h->UaN = p; break; case 18: h->WaN = p; break; case 19: h->W3N = p; break; case 20: h->WpN = p; break; case 21: h->WoN = p; break;
        }
    });
    h->bound = true; h->fresh = false; h->train_valid = false;
    return S2VT_OK;
}
extern "C" int s2vt_att_load_param(s2vt_att_handle* h, const char* tf_name, const float* src_host, const int64_t* shape, int ndim, s2vt_stream st) {
    if (!h || !tf_name || !src_host || !shape) return S2VT_EINVAL;
    if (!h->bound) return h->fail(S2VT_ESTATE, "s2vt_att_bind first");
    for (auto& v : h->vars) {
        bool match = v.name == tf_name;
        for (auto& a : v.aliases) match = match || a == tf_name;
        if (!match) continue;
        const int nd = v.cols ? 2 : 1;
        if (ndim != nd || shape[0] != v.rows || (nd == 2 && shape[1] != v.cols)) return h->fail(S2VT_ESHAPE, "%s: shape differs", tf_name);
        if (cudaMemcpyAsync(h->params + v.off, src_host, v.count() * 4, cudaMemcpyHostToDevice, (cudaStream_t)st) != cudaSuccess) return h->fail(S2VT_ECUDA, "copy of %s failed", tf_name);
        h->fresh = false;
        return S2VT_OK;
    }
    return h->fail(S2VT_ENOTFOUND, "no variable named %s", tf_name);
}

#define ACHK(h) do { (h)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return (h)->fail(S2VT_ECUDA, "launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)
#define AMEMSET(h, p, bytes, st) do { if (cudaMemsetAsync((p), 0, (bytes), (st)) != cudaSuccess) return (h)->fail(S2VT_ECUDA, "memset failed (%s:%d)", __FILE__, __LINE__); } while (0)

template <typename T>
static int att_refresh_impl(s2vt_att_handle* h, cudaStream_t st) {
    const int H = h->H, Hp = h->Hp, Dp = h->Dp, Vp = h->Vp;
    AMEMSET(h, h->WeT, (size_t)((char*)h->bo_p + (size_t)Vp * 4 - (char*)h->WeT), st);
    auto pk = [&](int var, void* dst, int ldd, int Rb, int Cb, int Cbp, int transpose) {
        const AttVar& v = h->vars[var];
        att_pack_kernel<T><<<592, 256, 0, st>>>(h->P_(var), (int)v.cols, (int)v.rows, (int)v.cols, (T*)dst, ldd, Rb, Cb, Hp, Cbp, transpose);
    };
    pk(h->iWe, h->WeT, Dp, h->D, H, Hp, 1); ACHK(h);               // [D, H]  -> [Hp, Dp]
    pk(h->iUa, h->UaT, Hp, H, H, Hp, 1); ACHK(h);
    pk(h->iWa, h->WaT, Hp, H, H, Hp, 1); ACHK(h);
    pk(h->iW3, h->W3T, 3 * Hp, H, H, Hp, 1); ACHK(h);              // [3H, 4H] -> [4Hp, 3Hp], input blocks and gate blocks padded
    pk(h->iWp, h->WpT, 3 * Hp, H, H, Hp, 1); ACHK(h);              // [3H, H]  -> [Hp, 3Hp]
    pk(h->iWo, h->WoT, Hp, H, h->V, Vp, 1); ACHK(h);               // [H, V]   -> [Vp, Hp]
    pk(h->iUa, h->UaN, Hp, H, H, Hp, 0); ACHK(h);                  // backward operands keep the TF orientation
    pk(h->iWa, h->WaN, Hp, H, H, Hp, 0); ACHK(h);
    pk(h->iW3, h->W3N, 4 * Hp, H, H, Hp, 0); ACHK(h);
    pk(h->iWp, h->WpN, Hp, H, H, Hp, 0); ACHK(h);
    pk(h->iWo, h->WoN, Vp, H, h->V, Vp, 0); ACHK(h);
    auto pv = [&](int var, float* dst, int Cb, int Cbp) { att_pack_vec_kernel<<<16, 256, 0, st>>>(h->P_(var), (int)h->vars[var].count(), dst, Cb, Cbp); };
    pv(h->ibe, h->be_p, H, Hp); ACHK(h); pv(h->iba, h->ba_p, H, Hp); ACHK(h); pv(h->iw, h->w_p, H, Hp); ACHK(h);
    pv(h->ib3, h->b3_p, H, Hp); ACHK(h); pv(h->ibp, h->bp_p, H, Hp); ACHK(h); pv(h->ibo, h->bo_p, h->V, Vp); ACHK(h);
    h->fresh = true; h->train_valid = false;
    return S2VT_OK;
}
extern "C" int s2vt_att_refresh(s2vt_att_handle* h, s2vt_stream st) {
    if (!h) return S2VT_EINVAL;
    if (!h->bound) return h->fail(S2VT_ESTATE, "s2vt_att_bind first");
    return h->cfg.precision == S2VT_PREC_BF16 ? att_refresh_impl<bf16>(h, (cudaStream_t)st) : att_refresh_impl<float>(h, (cudaStream_t)st);
}

// C = A . B^T (+ bias) (+ old C) through EpiStore: tcgen05 tiles for bf16 (128x256 / 128x128 batched, 128x32 for <= 128 rows), SIMT for fp32
template <typename T>
static int att_gemm(s2vt_att_handle* h, cudaStream_t st, const void* A, int lda, const void* B, int ldb, int M, int N, int K, float* outF, void* outT, int ldo,
                    const float* bias, int accumulate = 0) {
    typename EpiStore<T>::Params ep = {outF, (T*)outT, ldo, bias, M, accumulate};
    h->launches++;
    cudaError_t e;
    if constexpr (std::is_same<T, bf16>::value) {
        if (!h->maps) h->maps = new tc::MapCache();
        if (M > 128) e = N % 256 == 0 ? tc::launch<256, EpiStore<bf16>>(*h->maps, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false)
                                      : tc::launch<128, EpiStore<bf16>>(*h->maps, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false);
        // per-step GEMMs: B is always a packed weight matrix (written only by s2vt_att_refresh), so they run as programmatic dependents --
        // prologue, TMEM allocation and the first weight tiles overlap the element-wise kernel before them (which triggers early)
        else e = tc::launch<32, EpiStore<bf16>>(*h->maps, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true);
    } else {
        if (M > 64) e = launch_gemm<float, CfgBig, EpiStore<float>>(st, (const float*)A, lda, (const float*)B, ldb, M, N, K, ep);
        else e = launch_gemm<float, CfgStep, EpiStore<float>>(st, (const float*)A, lda, (const float*)B, ldb, M, N, K, ep);
    }
    if (e != cudaSuccess) return h->fail(S2VT_ECUDA, "GEMM [%d x %d x %d] failed: %s", M, N, K, cudaGetErrorString(e));
    return S2VT_OK;
}

// Weight gradient grad(var) += X^T . Y over `rows` rows (X [rows, ldx]: Mp padded features, Y [rows, ldy]: Np padded features) through the
// padded fp32 scratch tile.  bf16: MN-major tcgen05 operands read X and Y as they lie; fp32: explicit transposes + SIMT GEMM.
template <typename T>
static int att_wgrad(s2vt_att_handle* h, cudaStream_t st, const AttWork& w, const void* X, int ldx, int Mp, const void* Y, int ldy, int Np, int rows, int var,
                     int Rb, int Cb, int Cbp) {
    typename EpiStore<T>::Params ep = {w.tmpW, nullptr, Np, nullptr, Mp, 0};
    h->launches++;
    cudaError_t e;
    if constexpr (std::is_same<T, bf16>::value) {
        if (!h->maps) h->maps = new tc::MapCache();
        e = Np % 256 == 0 ? tc::launch<256, EpiStore<bf16>, 1, 1, 1, true>(*h->maps, st, (const bf16*)X, ldx, (const bf16*)Y, ldy, Mp, Np, rows, ep, false)
                          : tc::launch<128, EpiStore<bf16>, 1, 1, 1, true>(*h->maps, st, (const bf16*)X, ldx, (const bf16*)Y, ldy, Mp, Np, rows, ep, false);
    } else {
        const int Rp = ru(rows, 16);
        AMEMSET(h, w.tA, (size_t)Mp * Rp * 4, st); AMEMSET(h, w.tB, (size_t)Np * Rp * 4, st);
        att_transpose_kernel<float><<<592, 256, 0, st>>>((const float*)X, ldx, rows, Mp, (float*)w.tA, Rp); ACHK(h);
        att_transpose_kernel<float><<<592, 256, 0, st>>>((const float*)Y, ldy, rows, Np, (float*)w.tB, Rp); ACHK(h);
        e = launch_gemm<float, CfgBig, EpiStore<float>>(st, (const float*)w.tA, Rp, (const float*)w.tB, Rp, Mp, Np, Rp, ep);
    }
    if (e != cudaSuccess) return h->fail(S2VT_ECUDA, "weight-gradient GEMM [%d x %d x %d] failed: %s", Mp, Np, rows, cudaGetErrorString(e));
    const AttVar& v = h->vars[var];
    att_unpack_add_kernel<<<592, 256, 0, st>>>(w.tmpW, Np, (int)v.rows, (int)v.cols, h->grads + v.off, Rb, Cb, h->Hp, Cbp); ACHK(h);
    return S2VT_OK;
}

static int att_ready(s2vt_att_handle* h) {
    if (!h) return S2VT_EINVAL;
    if (!h->bound) return h->fail(S2VT_ESTATE, "s2vt_att_bind first");
    if (!h->fresh) return h->fail(S2VT_ESTATE, "parameters changed: call s2vt_att_refresh");
    return S2VT_OK;
}

// mode 0: greedy decode (ids_out, alphas_out).  mode 1: teacher-forced loss (captions, mask -> loss_out[2], logits_out).
// mode 2: mode 1 with every per-step operand stashed, followed by the backward pass into h->grads.
template <typename T>
static int att_run(s2vt_att_handle* h, cudaStream_t st, int mode, const float* video, int B, const int32_t* captions, const float* mask, uint64_t drop_seed,
                   uint32_t row_base, int32_t* ids_out, float* alphas_out, float* loss_out, float* logits_out) {
    const int n = h->n, H = h->H, Hp = h->Hp, Dp = h->Dp, Vp = h->Vp, R = B, Tc = h->Tc;
    const bool train = mode == 2;
    Arena a(h->ws, h->ws_bytes);
    AttWork w;
    att_plan(h, a, B, train, w);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace holds %zu bytes, %zu needed for %d videos%s", h->ws_bytes, a.used, B, train ? " (training)" : "");
    const size_t F = (size_t)B * n, e = h->esz, S = train ? (size_t)Tc : 1;
    // step-t views of the (possibly stashed) per-step buffers; slot S of hq / x3 / x4 / c takes the writes "for the step after the last"
    auto st_ = [&](int t) { return train ? (size_t)t : (size_t)0; };
    auto nx_ = [&](int t) { return train ? (size_t)t + 1 : (size_t)0; };
    auto HQ = [&](size_t s) { return (char*)w.hq + s * R * Hp * e; };
    auto X3 = [&](size_t s) { return (char*)w.x3 + s * R * 3 * Hp * e; };
    auto X4 = [&](size_t s) { return (char*)w.x4 + s * R * 3 * Hp * e; };
    auto OT = [&](size_t s) { return (char*)w.oT + s * R * Hp * e; };
    auto Cc = [&](size_t s) { return w.c + s * R * Hp; };
    // zero state: c, h_prev (hq), both concatenated operands (current_embed = 0 at step 0, padded lanes stay 0)
    AMEMSET(h, w.hq, (size_t)((char*)w.oT - (char*)w.hq), st);
    AMEMSET(h, w.c, (S + 1) * R * Hp * 4, st);
    AMEMSET(h, w.acc, 32, st);
    att_convert_video_kernel<T><<<592, 256, 0, st>>>(video, F, h->D, Dp, (T*)w.videoT); ACHK(h);
    ATRY((att_gemm<T>(h, st, w.videoT, Dp, h->WeT, Dp, (int)F, Hp, Dp, w.emb, w.embT, Hp, h->be_p)));            // image_emb (:95-96)
    ATRY((att_gemm<T>(h, st, w.embT, Hp, h->UaT, Hp, (int)F, Hp, Hp, w.part, nullptr, Hp, h->ba_p)));           // image_part (:107)
    const float keep = mode >= 1 ? h->cfg.dropout_keep : 1.0f;
    const int egrid = (int)std::min<size_t>(((size_t)R * Hp + 255) / 256, 1184);
    const int ggrid = (int)(((size_t)R * H + 255) / 256);
    for (int t = 0; t < Tc; ++t) {
        const size_t s = st_(t), sn = nx_(t);
        float* q = w.q + s * R * Hp; float* g = w.g + s * R * 4 * Hp; float* logits = w.logits + s * R * Vp;
        float* hinge = w.hinge + s * R; float* alph = w.alph + s * R * n;
        ATRY((att_gemm<T>(h, st, HQ(s), Hp, h->WaT, Hp, R, Hp, Hp, q, nullptr, Hp, nullptr)));                    // h_prev . Wa (:113)
        float* al = alphas_out ? alphas_out + (size_t)t * n * R : nullptr;
        att_attend_kernel<T><<<R, 256, 0, st>>>(q, w.part, w.emb, h->w_p, B, n, Hp, (T*)X3(s), (T*)X4(s), al, R, h->cfg.hinge_m, h->cfg.reg_frames, hinge,
                                                 train ? alph : nullptr); ACHK(h);
        ATRY((att_gemm<T>(h, st, X3(s), 3 * Hp, h->W3T, 3 * Hp, R, 4 * Hp, 3 * Hp, g, nullptr, 4 * Hp, h->b3_p)));  // LSTM3 pre-activations (:131)
        att_cell_kernel<T><<<egrid, 256, 0, st>>>(g, Cc(s), Cc(sn), R, Hp, (T*)X3(sn), (T*)X4(s), (T*)HQ(sn), drop_seed, (uint32_t)t, row_base, keep); ACHK(h);
        ATRY((att_gemm<T>(h, st, X4(s), 3 * Hp, h->WpT, 3 * Hp, R, Hp, 3 * Hp, w.o, nullptr, Hp, h->bp_p)));        // head (:134)
        att_tanh_kernel<T><<<egrid, 256, 0, st>>>(w.o, (size_t)R * Hp, (T*)OT(s), train ? w.o2F + s * R * Hp : nullptr); ACHK(h);
        ATRY((att_gemm<T>(h, st, OT(s), Hp, h->WoT, Hp, R, Vp, Hp, logits, nullptr, Vp, h->bo_p)));               // logit_words (:143)
        if (logits_out && cudaMemcpy2DAsync(logits_out + (size_t)t * R * h->V, (size_t)h->V * 4, logits, (size_t)Vp * 4, (size_t)h->V * 4, R, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return h->fail(S2VT_ECUDA, "logits copy failed");
        if (mode == 0) {
            att_argmax_kernel<<<R, 256, 0, st>>>(logits, Vp, h->V, w.tok, ids_out, Tc, t); ACHK(h);
            att_gather_kernel<T><<<ggrid, 256, 0, st>>>(h->P_(h->iWemb), H, Hp, w.tok, 1, 0, R, (T*)X3(sn), (T*)X4(sn)); ACHK(h);
        } else {
            att_ce_kernel<<<R, 256, 0, st>>>(logits, Vp, h->V, captions, mask, Tc, t, hinge, h->cfg.hinge_beta, w.acc); ACHK(h);
            att_gather_kernel<T><<<ggrid, 256, 0, st>>>(h->P_(h->iWemb), H, Hp, captions, Tc, t, R, (T*)X3(sn), (T*)X4(sn)); ACHK(h);
        }
    }
    if (mode >= 1) { att_loss_final_kernel<<<1, 1, 0, st>>>(w.acc, loss_out, train ? h->grads + h->P : nullptr); ACHK(h); }
    if (!train) return S2VT_OK;

    // ---------------- backward ----------------
    const int TR = Tc * R;
    AMEMSET(h, h->grads, h->P * 4, st);
    AMEMSET(h, h->sq, 32, st);
    AMEMSET(h, w.dc, (size_t)R * Hp * 4, st);
    AMEMSET(h, w.d_part, (size_t)((char*)w.tmpW - (char*)w.d_part), st);        // d_part, d_emb, dw
    // everything after the recurrence, batched over the T_c steps
    att_dlogits_kernel<T><<<TR, 256, 0, st>>>(w.logits, Vp, h->V, captions, mask, R, Tc, w.acc, (T*)w.dz); ACHK(h);
    ATRY((att_wgrad<T>(h, st, w, w.oT, Hp, Hp, w.dz, Vp, Vp, TR, h->iWo, H, h->V, Vp)));                         // dWo = o2^T dz
    att_colsum_kernel<T><<<(h->V + 255) / 256, 256, 0, st>>>((const T*)w.dz, Vp, TR, h->V, h->V, Vp, h->grads + h->vars[h->ibo].off); ACHK(h);
    ATRY((att_gemm<T>(h, st, w.dz, Vp, h->WoN, Vp, TR, Hp, Vp, w.do2, nullptr, Hp, nullptr)));                   // do2 = dz . Wo^T
    att_du_kernel<T><<<egrid, 256, 0, st>>>(w.do2, w.o2F, (size_t)TR * Hp, (T*)w.du); ACHK(h);
    ATRY((att_wgrad<T>(h, st, w, w.x4, 3 * Hp, 3 * Hp, w.du, Hp, Hp, TR, h->iWp, H, H, Hp)));                    // dWp = x4^T du
    att_colsum_kernel<T><<<(H + 255) / 256, 256, 0, st>>>((const T*)w.du, Hp, TR, H, H, Hp, h->grads + h->vars[h->ibp].off); ACHK(h);
    ATRY((att_gemm<T>(h, st, w.du, Hp, h->WpN, Hp, TR, 3 * Hp, Hp, w.dx4, nullptr, 3 * Hp, nullptr)));           // dx4 = du . Wp^T
    // reverse time loop
    const size_t smem = ((size_t)Hp + n) * 4;
    for (int t = Tc - 1; t >= 0; --t) {
        float* dx3 = (t & 1) ? w.dx3a : w.dx3b;            // this step's dx3; the other buffer holds step t+1's
        const float* dx3_next = t + 1 < Tc ? ((t & 1) ? w.dx3b : w.dx3a) : nullptr;
        const float* dx4 = w.dx4 + (size_t)t * R * 3 * Hp;
        T* dg = (T*)w.dg + (size_t)t * R * 4 * Hp;
        att_cell_bwd_kernel<T><<<egrid, 256, 0, st>>>(dx4, t + 1 < Tc ? w.dhq : nullptr, dx3_next, w.g + (size_t)t * R * 4 * Hp, Cc(t), Cc(t + 1), w.dc, R, Hp, dg,
                                                      drop_seed, (uint32_t)t, row_base, keep); ACHK(h);
        ATRY((att_gemm<T>(h, st, dg, 4 * Hp, h->W3N, 4 * Hp, R, 3 * Hp, 4 * Hp, dx3, nullptr, 3 * Hp, nullptr)));  // dx3 = dg . W3^T
        T* dq = (T*)w.dq + (size_t)t * R * Hp;
        att_attend_bwd_kernel<T><<<R, 256, smem, st>>>(dx4, dx3, w.q + (size_t)t * R * Hp, w.part, w.emb, h->w_p, w.alph + (size_t)t * R * n, w.hinge + (size_t)t * R, mask,
                                                       Tc, t, w.acc, h->cfg.hinge_beta, h->cfg.reg_frames, B, n, Hp, dq, w.d_part, w.d_emb, w.dw); ACHK(h);
        if (t > 0) {
            att_scatter_emb_kernel<<<std::min(ggrid, 592), 256, 0, st>>>(dx4, dx3, captions, Tc, t, R, H, Hp, h->grads + h->vars[h->iWemb].off, h->sq); ACHK(h);
            ATRY((att_gemm<T>(h, st, dq, Hp, h->WaN, Hp, R, Hp, Hp, w.dhq, nullptr, Hp, nullptr)));               // d h_prev = dq . Wa^T
        }
    }
    // weight gradients over the stashed per-step operands
    ATRY((att_wgrad<T>(h, st, w, w.x3, 3 * Hp, 3 * Hp, w.dg, 4 * Hp, 4 * Hp, TR, h->iW3, H, H, Hp)));            // dW3 = x3^T dg
    att_colsum_kernel<T><<<(4 * H + 255) / 256, 256, 0, st>>>((const T*)w.dg, 4 * Hp, TR, 4 * H, H, Hp, h->grads + h->vars[h->ib3].off); ACHK(h);
    ATRY((att_wgrad<T>(h, st, w, w.hq, Hp, Hp, w.dq, Hp, Hp, TR, h->iWa, H, H, Hp)));                            // dWa = h_prev^T dq
    att_pack_vec_kernel<<<4, 256, 0, st>>>(w.dw, H, h->grads + h->vars[h->iw].off, H, H); ACHK(h);               // dw (grads were zeroed: plain copy)
    // frames: image_part = emb . Ua + ba ; image_emb = X . We + be
    att_cast_kernel<T><<<592, 256, 0, st>>>(w.d_part, F * Hp, (T*)w.d_partT); ACHK(h);
    ATRY((att_wgrad<T>(h, st, w, w.embT, Hp, Hp, w.d_partT, Hp, Hp, (int)F, h->iUa, H, H, Hp)));                 // dUa = emb^T d_part
    att_colsum_kernel<float><<<(H + 255) / 256, 256, 0, st>>>(w.d_part, Hp, (int)F, H, H, Hp, h->grads + h->vars[h->iba].off); ACHK(h);
    ATRY((att_gemm<T>(h, st, w.d_partT, Hp, h->UaN, Hp, (int)F, Hp, Hp, w.d_emb, nullptr, Hp, nullptr, 1)));     // d_emb += d_part . Ua^T
    att_cast_kernel<T><<<592, 256, 0, st>>>(w.d_emb, F * Hp, (T*)w.d_embT); ACHK(h);
    ATRY((att_wgrad<T>(h, st, w, w.videoT, Dp, Dp, w.d_embT, Hp, Hp, (int)F, h->iWe, h->D, H, Hp)));            // dWe = X^T d_emb ([D, H]: one row block)
    att_colsum_kernel<float><<<(H + 255) / 256, 256, 0, st>>>(w.d_emb, Hp, (int)F, H, H, Hp, h->grads + h->vars[h->ibe].off); ACHK(h);
    h->train_valid = true; h->train_B = B;
    return S2VT_OK;
}

extern "C" int s2vt_att_greedy(s2vt_att_handle* h, const float* video, int B, int32_t* ids_out, float* alphas_out, s2vt_stream st) {
    ATRY(att_ready(h));
    if (!video || !ids_out || B <= 0) return h->fail(S2VT_EINVAL, "bad argument");
    return h->cfg.precision == S2VT_PREC_BF16 ? att_run<bf16>(h, (cudaStream_t)st, 0, video, B, nullptr, nullptr, 0, 0, ids_out, alphas_out, nullptr, nullptr)
                                              : att_run<float>(h, (cudaStream_t)st, 0, video, B, nullptr, nullptr, 0, 0, ids_out, alphas_out, nullptr, nullptr);
}
extern "C" int s2vt_att_xe_loss(s2vt_att_handle* h, const float* video, int B, const int32_t* captions, const float* mask, uint64_t drop_seed, uint32_t row_base,
                                float* loss_out, float* logits_out, s2vt_stream st) {
    ATRY(att_ready(h));
    if (!video || !captions || !mask || !loss_out || B <= 0) return h->fail(S2VT_EINVAL, "bad argument");
    return h->cfg.precision == S2VT_PREC_BF16 ? att_run<bf16>(h, (cudaStream_t)st, 1, video, B, captions, mask, drop_seed, row_base, nullptr, nullptr, loss_out, logits_out)
                                              : att_run<float>(h, (cudaStream_t)st, 1, video, B, captions, mask, drop_seed, row_base, nullptr, nullptr, loss_out, logits_out);
}
extern "C" int s2vt_att_xe_backward(s2vt_att_handle* h, const float* video, int B, const int32_t* captions, const float* mask, uint64_t drop_seed, uint32_t row_base,
                                    float* loss_out, s2vt_stream st) {
    ATRY(att_ready(h));
    if (!video || !captions || !mask || !loss_out || B <= 0) return h->fail(S2VT_EINVAL, "bad argument");
    return h->cfg.precision == S2VT_PREC_BF16 ? att_run<bf16>(h, (cudaStream_t)st, 2, video, B, captions, mask, drop_seed, row_base, nullptr, nullptr, loss_out, nullptr)
                                              : att_run<float>(h, (cudaStream_t)st, 2, video, B, captions, mask, drop_seed, row_base, nullptr, nullptr, loss_out, nullptr);
}
extern "C" int s2vt_att_optimizer_step(s2vt_att_handle* h, float lr, float clip_norm, int64_t step, float* out, s2vt_stream st_) {
    ATRY(att_ready(h));
    if (!h->train_valid) return h->fail(S2VT_ESTATE, "s2vt_att_xe_backward first");
    if (step <= 0) return h->fail(S2VT_EINVAL, "Adam step counts from 1");
    cudaStream_t st = (cudaStream_t)st_;
    const AttVar& we = h->vars[h->iWemb];
    att_sumsq_kernel<<<592, 256, 0, st>>>(h->grads, h->P, h->sq); ACHK(h);
    att_sumsq_kernel<<<148, 256, 0, st>>>(h->grads + we.off, we.count(), h->sq + 1); ACHK(h);
    const double b1 = 0.9, b2 = 0.999;
    const float lr_t = (float)(lr * sqrt(1.0 - pow(b2, (double)step)) / (1.0 - pow(b1, (double)step)));
    att_adam_kernel<<<592, 256, 0, st>>>(h->params, h->grads, h->adam_m, h->adam_v, h->P, h->sq, clip_norm, lr_t, 0.9f, 0.999f, 1e-8f, out); ACHK(h);
    h->train_valid = false;
    return h->cfg.precision == S2VT_PREC_BF16 ? att_refresh_impl<bf16>(h, st) : att_refresh_impl<float>(h, st);
}
extern "C" long long s2vt_att_launch_count(const s2vt_att_handle* h) { return h ? h->launches : 0; }
