// Temporal-attention caption decoder (BASELINE config 3): the model of original_attention.py:54-251 --
// frame projection, additive attention over the frame embeddings, single LSTM3, tanh MLP head, vocabulary projection --
// as the greedy sampler (build_generator :155-199 / build_sampler :201-251, with the saved alphas) and the teacher-forced
// loss of build_model (:88-152, DropoutWrapper on the LSTM3 output, hinge regulariser on the first 8 alphas).
//
// Contractions run on the library's GEMM mainloops (tcgen05 / TMA / TMEM for bf16, SIMT FMA for the fp32 mode) with the
// EpiStore epilogue; the per-step glue (attention scores + softmax + weighted frame sum, LSTM cell, head non-linearity,
// arg-max, embedding gather, cross entropy) is the small kernels below.  Per decode step and row the attention reads the
// video's [n, H] `image_part` and `image_emb` tables (fp32) once: 2 n H 4 B, HBM/L2-bound.
// Layout: frames are video-major rows b * n + i (the reference's [n, b, h] transpose is only a view); LSTM3 gate columns
// keep the TF order g * Hp + u; the concatenated operands [atten | emb | h] and [out1 | atten | emb] are kept as
// three Hp-wide blocks of one row so each product is ONE GEMM with K = 3 Hp.
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
    float* params = nullptr;
    void *WeT, *UaT, *WaT, *W3T, *WpT, *WoT;
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
template <typename T>
__global__ void att_pack_kernel(const float* __restrict__ src, int lds, int R, int C, T* __restrict__ dst, int ldd, int Rb, int Cb, int Rbp, int Cbp) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * C; idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx / C), c = (int)(idx % C);
        dst[(size_t)((c / Cb) * Cbp + c % Cb) * ldd + (r / Rb) * Rbp + r % Rb] = from_f32<T>(src[(size_t)r * lds + c]);
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
                                                         float* __restrict__ alphas_out, int R, float m_hinge, int reg_frames, float* __restrict__ hinge) {
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
    for (int h = tid; h < Hp; h += blockDim.x) {
        float a = 0.f;
        for (int i = 0; i < n; ++i) a += e[i] * inv * emb[((size_t)v * n + i) * Hp + h];
        T t = from_f32<T>(a);
        x3[(size_t)r * 3 * Hp + h] = t;
        x4[(size_t)r * 3 * Hp + Hp + h] = t;
    }
}

// BasicLSTMCell on the pre-activations g [R, 4Hp] (TF gate order i, j, f, o in blocks of Hp, bias already added): state c in
// place, un-dropped h into block 2 of x3 (the cell's own recurrent input), DropoutWrapper output into hq (next step's attention
// query operand, h_prev = output1 :135) and block 0 of x4 (head operand).
template <typename T>
__global__ void att_cell_kernel(const float* __restrict__ g, float* __restrict__ c, int R, int Hp, T* __restrict__ x3, T* __restrict__ x4, T* __restrict__ hq,
                                unsigned long long seed, uint32_t step, uint32_t row_base, float keep) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)R * Hp; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / Hp), u = (int)(idx % Hp);
        const float* gr = g + (size_t)r * 4 * Hp;
        const float si = sigm<T>(gr[u]), tj = tanh_<T>(gr[Hp + u]), sf = sigm<T>(gr[2 * Hp + u] + 1.0f), so = sigm<T>(gr[3 * Hp + u]);
        const float cn = c[idx] * sf + si * tj;
        const float hn = tanh_<T>(cn) * so;
        c[idx] = cn;
        x3[(size_t)r * 3 * Hp + 2 * Hp + u] = from_f32<T>(hn);
        const float out = keep < 1.0f ? hn * dropout_mult(seed, S2VT_STREAM_DROP1, row_base + (uint32_t)r, step, (uint32_t)u, keep) : hn;
        hq[idx] = from_f32<T>(out);
        x4[(size_t)r * 3 * Hp + u] = from_f32<T>(out);
    }
}
template <typename T>
__global__ void att_tanh_kernel(const float* __restrict__ x, size_t count, T* __restrict__ out) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < count; idx += (size_t)gridDim.x * blockDim.x) out[idx] = from_f32<T>(tanh_<T>(x[idx]));
}
// tf.argmax(logit_words, 1) (:190): lowest index among equal maxima; also the id matrix column.
__global__ void __launch_bounds__(256) att_argmax_kernel(const float* __restrict__ logits, int ld, int V, int* __restrict__ tok, int* __restrict__ ids, int Tc, int t) {
    __shared__ ArgVal red[32];
    const int r = blockIdx.x;
    ArgVal best; best.v = -INFINITY; best.i = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += blockDim.x) { ArgVal c; c.v = logits[(size_t)r * ld + v]; c.i = v; best = argmax_op(best, c); }
    best = block_argmax(best, red);
    if (threadIdx.x == 0) { tok[r] = best.i; ids[(size_t)r * Tc + t] = best.i; }
}
// current_embed = Wemb[word] (:148, 192) into block 1 of x3 and block 2 of x4; word = tok[r] or caption[r, t].
template <typename T>
__global__ void att_gather_kernel(const float* __restrict__ Wemb, int H, int Hp, const int* __restrict__ tok, int tok_ld, int tok_col, int R, T* __restrict__ x3,
                                  T* __restrict__ x4) {
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
__global__ void att_loss_final_kernel(const double* __restrict__ acc, float* __restrict__ out) {
    out[0] = (float)(acc[0] / acc[2]);
    out[1] = (float)(acc[1] / acc[2]);
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
    assign(1, a.take<char>((size_t)Hp * Dp * e)); assign(2, a.take<char>((size_t)Hp * Hp * e)); assign(3, a.take<char>((size_t)Hp * Hp * e));
    assign(4, a.take<char>((size_t)4 * Hp * 3 * Hp * e)); assign(5, a.take<char>((size_t)Hp * 3 * Hp * e)); assign(6, a.take<char>((size_t)Vp * Hp * e));
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

// per-call scratch, R rows of B videos
struct AttWork {
    void *videoT, *embT, *hq, *x3, *x4, *oT;
    float *emb, *part, *q, *g, *c, *o, *logits, *hinge, *alph;
    int* tok; double* acc;
};
static void att_plan(const s2vt_att_handle* h, Arena& a, int B, int R, AttWork& w) {
    const size_t e = h->esz, F = (size_t)B * h->n;
    w.videoT = a.take<char>(F * h->Dp * e); w.embT = a.take<char>(F * h->Hp * e);
    w.emb = a.take<float>(F * h->Hp); w.part = a.take<float>(F * h->Hp);
    w.hq = a.take<char>((size_t)R * h->Hp * e); w.x3 = a.take<char>((size_t)R * 3 * h->Hp * e); w.x4 = a.take<char>((size_t)R * 3 * h->Hp * e);
    w.oT = a.take<char>((size_t)R * h->Hp * e);
    w.q = a.take<float>((size_t)R * h->Hp); w.g = a.take<float>((size_t)R * 4 * h->Hp); w.c = a.take<float>((size_t)R * h->Hp); w.o = a.take<float>((size_t)R * h->Hp);
    w.logits = a.take<float>((size_t)R * h->Vp); w.hinge = a.take<float>(R); w.alph = a.take<float>((size_t)R * h->n);
    w.tok = a.take<int>(R); w.acc = a.take<double>(4);
}
extern "C" size_t s2vt_att_workspace_bytes(const s2vt_att_handle* h, int n_videos, int n_rows) {
    if (!h || n_videos <= 0 || n_rows < n_videos) return 0;
    Arena a(nullptr, 0);
    AttWork w;
    att_plan(h, a, n_videos, n_rows, w);
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
        }
    });
    h->bound = true; h->fresh = false;
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

template <typename T>
static int att_refresh_impl(s2vt_att_handle* h, cudaStream_t st) {
    const int H = h->H, Hp = h->Hp, Dp = h->Dp;
    const size_t copies = (size_t)((char*)h->bo_p + (size_t)h->Vp * 4 - (char*)h->WeT);
    if (cudaMemsetAsync(h->WeT, 0, copies, st) != cudaSuccess) return h->fail(S2VT_ECUDA, "memset failed");
    auto pk = [&](int var, void* dst, int ldd, int Rb, int Cb, int Cbp) {
        const AttVar& v = h->vars[var];
        att_pack_kernel<T><<<592, 256, 0, st>>>(h->P_(var), (int)v.cols, (int)v.rows, (int)v.cols, (T*)dst, ldd, Rb, Cb, Hp, Cbp);
    };
    pk(h->iWe, h->WeT, Dp, h->D, H, Hp); ACHK(h);                  // [D, H]  -> [Hp, Dp]
    pk(h->iUa, h->UaT, Hp, H, H, Hp); ACHK(h);
    pk(h->iWa, h->WaT, Hp, H, H, Hp); ACHK(h);
    pk(h->iW3, h->W3T, 3 * Hp, H, H, Hp); ACHK(h);                 // [3H, 4H] -> [4Hp, 3Hp], input blocks and gate blocks padded
    pk(h->iWp, h->WpT, 3 * Hp, H, H, Hp); ACHK(h);                 // [3H, H]  -> [Hp, 3Hp]
    pk(h->iWo, h->WoT, Hp, H, h->V, h->Vp); ACHK(h);               // [H, V]   -> [Vp, Hp]
    auto pv = [&](int var, float* dst, int Cb, int Cbp) { att_pack_vec_kernel<<<16, 256, 0, st>>>(h->P_(var), (int)h->vars[var].count(), dst, Cb, Cbp); };
    pv(h->ibe, h->be_p, H, Hp); ACHK(h); pv(h->iba, h->ba_p, H, Hp); ACHK(h); pv(h->iw, h->w_p, H, Hp); ACHK(h);
    pv(h->ib3, h->b3_p, H, Hp); ACHK(h); pv(h->ibp, h->bp_p, H, Hp); ACHK(h); pv(h->ibo, h->bo_p, h->V, h->Vp); ACHK(h);
    h->fresh = true;
    return S2VT_OK;
}
extern "C" int s2vt_att_refresh(s2vt_att_handle* h, s2vt_stream st) {
    if (!h) return S2VT_EINVAL;
    if (!h->bound) return h->fail(S2VT_ESTATE, "s2vt_att_bind first");
    return h->cfg.precision == S2VT_PREC_BF16 ? att_refresh_impl<bf16>(h, (cudaStream_t)st) : att_refresh_impl<float>(h, (cudaStream_t)st);
}

// C = A . B^T (+ bias) through EpiStore: tcgen05 tiles for bf16 (128x256 / 128x128 batched, 128x32 for <= 128 rows), SIMT for fp32
template <typename T>
static int att_gemm(s2vt_att_handle* h, cudaStream_t st, const void* A, int lda, const void* B, int ldb, int M, int N, int K, float* outF, void* outT, int ldo,
                    const float* bias) {
    typename EpiStore<T>::Params ep = {outF, (T*)outT, ldo, bias, M, 0};
    h->launches++;
    cudaError_t e;
    if constexpr (std::is_same<T, bf16>::value) {
        if (!h->maps) h->maps = new tc::MapCache();
        if (M > 128) e = N % 256 == 0 ? tc::launch<256, EpiStore<bf16>>(*h->maps, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false)
                                      : tc::launch<128, EpiStore<bf16>>(*h->maps, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false);
        else e = tc::launch<32, EpiStore<bf16>>(*h->maps, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false);
    } else {
        if (M > 64) e = launch_gemm<float, CfgBig, EpiStore<float>>(st, (const float*)A, lda, (const float*)B, ldb, M, N, K, ep);
        else e = launch_gemm<float, CfgStep, EpiStore<float>>(st, (const float*)A, lda, (const float*)B, ldb, M, N, K, ep);
    }
    if (e != cudaSuccess) return h->fail(S2VT_ECUDA, "GEMM [%d x %d x %d] failed: %s", M, N, K, cudaGetErrorString(e));
    return S2VT_OK;
}

static int att_ready(s2vt_att_handle* h) {
    if (!h) return S2VT_EINVAL;
    if (!h->bound) return h->fail(S2VT_ESTATE, "s2vt_att_bind first");
    if (!h->fresh) return h->fail(S2VT_ESTATE, "parameters changed: call s2vt_att_refresh");
    return S2VT_OK;
}

// mode 0: greedy decode (ids_out, alphas_out).  mode 1: teacher-forced loss (captions, mask -> loss_out[2], logits_out).
template <typename T>
static int att_run(s2vt_att_handle* h, cudaStream_t st, int mode, const float* video, int B, const int32_t* captions, const float* mask, uint64_t drop_seed,
                   uint32_t row_base, int32_t* ids_out, float* alphas_out, float* loss_out, float* logits_out) {
    const int n = h->n, Hp = h->Hp, Dp = h->Dp, Vp = h->Vp, R = B, Tc = h->Tc;
    Arena a(h->ws, h->ws_bytes);
    AttWork w;
    att_plan(h, a, B, R, w);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace holds %zu bytes, %zu needed for %d videos", h->ws_bytes, a.used, B);
    const size_t F = (size_t)B * n, e = h->esz;
    // zero state: c, h_prev (hq), both concatenated operands (current_embed = 0 at step 0, padded lanes stay 0)
    if (cudaMemsetAsync(w.hq, 0, (size_t)((char*)w.oT - (char*)w.hq), st) != cudaSuccess || cudaMemsetAsync(w.c, 0, (size_t)R * Hp * 4, st) != cudaSuccess ||
        cudaMemsetAsync(w.acc, 0, 32, st) != cudaSuccess)
        return h->fail(S2VT_ECUDA, "memset failed");
    att_convert_video_kernel<T><<<592, 256, 0, st>>>(video, F, h->D, Dp, (T*)w.videoT); ACHK(h);
    ATRY((att_gemm<T>(h, st, w.videoT, Dp, h->WeT, Dp, (int)F, Hp, Dp, w.emb, w.embT, Hp, h->be_p)));            // image_emb (:95-96)
    ATRY((att_gemm<T>(h, st, w.embT, Hp, h->UaT, Hp, (int)F, Hp, Hp, w.part, nullptr, Hp, h->ba_p)));           // image_part (:107)
    const float keep = mode == 1 ? h->cfg.dropout_keep : 1.0f;
    const int egrid = (int)((((size_t)R * Hp + 255) / 256) < 1184 ? (((size_t)R * Hp + 255) / 256) : 1184);
    for (int t = 0; t < Tc; ++t) {
        ATRY((att_gemm<T>(h, st, w.hq, Hp, h->WaT, Hp, R, Hp, Hp, w.q, nullptr, Hp, nullptr)));                   // h_prev . Wa (:113)
        float* al = alphas_out ? alphas_out + (size_t)t * n * R : nullptr;
        att_attend_kernel<T><<<R, 256, 0, st>>>(w.q, w.part, w.emb, h->w_p, B, n, Hp, (T*)w.x3, (T*)w.x4, al, R, h->cfg.hinge_m, h->cfg.reg_frames, w.hinge); ACHK(h);
        ATRY((att_gemm<T>(h, st, w.x3, 3 * Hp, h->W3T, 3 * Hp, R, 4 * Hp, 3 * Hp, w.g, nullptr, 4 * Hp, h->b3_p)));   // LSTM3 pre-activations (:131)
        att_cell_kernel<T><<<egrid, 256, 0, st>>>(w.g, w.c, R, Hp, (T*)w.x3, (T*)w.x4, (T*)w.hq, drop_seed, (uint32_t)t, row_base, keep); ACHK(h);
        ATRY((att_gemm<T>(h, st, w.x4, 3 * Hp, h->WpT, 3 * Hp, R, Hp, 3 * Hp, w.o, nullptr, Hp, h->bp_p)));           // head (:134)
        att_tanh_kernel<T><<<egrid, 256, 0, st>>>(w.o, (size_t)R * Hp, (T*)w.oT); ACHK(h);
        ATRY((att_gemm<T>(h, st, w.oT, Hp, h->WoT, Hp, R, Vp, Hp, w.logits, nullptr, Vp, h->bo_p)));                 // logit_words (:143)
        if (logits_out && cudaMemcpy2DAsync(logits_out + (size_t)t * R * h->V, (size_t)h->V * 4, w.logits, (size_t)Vp * 4, (size_t)h->V * 4, R, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return h->fail(S2VT_ECUDA, "logits copy failed");
        const int ggrid = (int)(((size_t)R * h->H + 255) / 256);
        if (mode == 0) {
            att_argmax_kernel<<<R, 256, 0, st>>>(w.logits, Vp, h->V, w.tok, ids_out, Tc, t); ACHK(h);
            att_gather_kernel<T><<<ggrid, 256, 0, st>>>(h->P_(h->iWemb), h->H, Hp, w.tok, 1, 0, R, (T*)w.x3, (T*)w.x4); ACHK(h);
        } else {
            att_ce_kernel<<<R, 256, 0, st>>>(w.logits, Vp, h->V, captions, mask, Tc, t, w.hinge, h->cfg.hinge_beta, w.acc); ACHK(h);
            att_gather_kernel<T><<<ggrid, 256, 0, st>>>(h->P_(h->iWemb), h->H, Hp, captions, Tc, t, R, (T*)w.x3, (T*)w.x4); ACHK(h);
        }
    }
    if (mode == 1) { att_loss_final_kernel<<<1, 1, 0, st>>>(w.acc, loss_out); ACHK(h); }
    (void)e;
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
extern "C" long long s2vt_att_launch_count(const s2vt_att_handle* h) { return h ? h->launches : 0; }
