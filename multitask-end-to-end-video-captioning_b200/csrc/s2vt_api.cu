// libs2vt_b200.so -- model entry points of the C ABI declared in include/s2vt.h.
//
// Data flow (one REINFORCE iteration; B videos, K samples, N = K*B rows, T = T_v + T_c steps):
//   frontend : video -> Xc (time-major, compute dtype) -> img = Xc.We + be -> G1x = img.W1x            [batched GEMMs]
//   LSTM1    : T recurrent steps over B rows only -- LSTM1 never sees a word (decoder input is `padding`,
//              reinforcement_multisampling_tf_s2vt.py:327-328), so it is shared by every sample of a video
//   G2x      : (dropout(h1)) . W2[out1 rows]  for all T steps at once                                   [batched GEMM]
//   LSTM2    : T recurrent steps, per row: h2.W2[h rows] + G2x + Etab[prev word] (Etab = Wemb.W2[emb rows])
//   logits   : out2 . embed_word_W + b                      [per step in rollouts, batched when teacher forced]
//   backward : mirror image; weight gradients as batched GEMMs over the stashed per-step gate gradients.
#include "engine.cuh"
#include <type_traits>

#include "gemm.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_chain.cuh"
#include "gemm_tcgen05_ws2.cuh"
#include "gemm_tcgen05_pair.cuh"
#include "gemm_tcgen05_persist.cuh"
#include "kernels.cuh"
#include "peer.cuh"

template <typename T, typename F>
__global__ void lstm_bwd_elem_kernel(LstmBwdArgs a, T* dg_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.M * a.Hp) return;
    lstm_bwd_unit<T, F>(a, dg_out, idx / a.Hp, idx % a.Hp, 0.f);
}

// out[r, :] = in[r % B, :]  (replicate the per-video encoder state over the K+1 decode rows)
template <typename U>
__global__ void tile_rows_kernel(const U* __restrict__ in, int B, int R, int ld, U* __restrict__ out) {
    int r = blockIdx.x;
    const U* s = in + (size_t)(r % B) * ld;
    U* d = out + (size_t)r * ld;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) d[c] = s[c];
}

// mean over frames of the fp32 feed: pooled[b, d] = mean_t video[b, t, d]  -> compute dtype [B, Dp]
template <typename T>
__global__ void mean_frames_kernel(const float* __restrict__ video, int Tv, int D, int Dp, T* __restrict__ out) {
    int b = blockIdx.x;
    for (int d = threadIdx.x; d < Dp; d += blockDim.x) {
        float acc = 0.f;
        if (d < D) for (int t = 0; t < Tv; ++t) acc += video[((size_t)b * Tv + t) * D + d];
        out[(size_t)b * Dp + d] = from_f32<T>(acc / (float)Tv);
    }
}

// sigmoid cross entropy with logits (tf.nn.sigmoid_cross_entropy_with_logits [lib]) / (A*B); dz in compute dtype
template <typename T>
__global__ void sigmoid_ce_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ y, int B, int A, int Ap, float scale, T* __restrict__ dz,
                                  float* __restrict__ loss) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int idx = threadIdx.x; idx < B * Ap; idx += blockDim.x) {
        int b = idx / Ap, a = idx % Ap;
        float d = 0.f;
        if (a < A) {
            float zz = z[(size_t)b * ldz + a], yy = y[(size_t)b * A + a];
            acc += fmaxf(zz, 0.f) - zz * yy + log1pf(expf(-fabsf(zz)));
            d = scale * (sigmoidf_(zz) - yy) / (float)(A * B);
        }
        dz[(size_t)b * Ap + a] = from_f32<T>(d);
    }
    acc = block_reduce(acc, [](float a, float b) { return a + b; }, red);
    if (threadIdx.x == 0) loss[0] = acc / (float)(A * B);
}

__global__ void add_scaled_scalar_kernel(float* dst, const float* src, float scale) { dst[0] += scale * src[0]; }
__global__ void copy_slice_norm_kernel(const float* src, double* dst) { dst[0] = (double)src[0]; }
__global__ void set_norm_kernel(float* nrm, const float* local_sum, float override_) { nrm[0] = override_ > 0.f ? override_ : local_sum[0]; }
__global__ void xe_total_kernel(float* out, const double* sq, float decay, float scale) {
    float wd = decay * 0.5f * (float)sq[3];
    out[1] = wd;
    out[0] += scale * wd;
}

// =================================================================================================================
// handle / memory
// =================================================================================================================
static void add_var(s2vt_handle* h, const char* name, std::vector<std::string> aliases, int64_t rows, int64_t cols) {
    Var v; v.name = name; v.aliases = aliases; v.rows = rows; v.cols = cols; v.off = h->P;
    h->P += ru64(v.count(), 4);
    h->vars.push_back(v);
}

template <typename F> static void layout_state(s2vt_handle* h, Arena& a, F assign) {
    const size_t e = h->esz;
    const int Dp = h->Dp, Ep = h->Ep, Hp = h->Hp, Vp = h->Vp, Gp = h->Gp, Ap = h->Ap;
    float* params = a.take<float>(h->P);
    float* grads = a.take<float>(h->P + 8);
    float* m = a.take<float>(h->P);
    float* v = a.take<float>(h->P);
    double* sq = a.take<double>(8);
    float* scal = a.take<float>(64 + 64);   // [64..127]: unsigned grid-barrier counters
    size_t copies_begin = a.used;
    char* WeT = a.take<char>((size_t)Ep * Dp * e);
    char* W1xT = a.take<char>((size_t)Gp * Ep * e);
    char* W1hT = a.take<char>((size_t)Gp * Hp * e);
    char* W1h = a.take<char>((size_t)Hp * Gp * e);
    char* W1x = a.take<char>((size_t)Ep * Gp * e);
    char* W2xT = a.take<char>((size_t)Gp * Hp * e);
    char* W2x = a.take<char>((size_t)Hp * Gp * e);
    char* W2eT = a.take<char>((size_t)Gp * Ep * e);
    char* W2e = a.take<char>((size_t)Ep * Gp * e);
    char* W2hT = a.take<char>((size_t)Gp * Hp * e);
    char* W2h = a.take<char>((size_t)Hp * Gp * e);
    char* WoT = a.take<char>((size_t)Vp * Hp * e);
    char* Wo = a.take<char>((size_t)Hp * Vp * e);
    char* WembC = a.take<char>((size_t)Vp * Ep * e);
    char* attrWT = a.take<char>((size_t)(Ap ? Ap : 1) * Dp * e);
    float* be_p = a.take<float>(Ep);
    float* b1_p = a.take<float>(Gp);
    float* b2_p = a.take<float>(Gp);
    float* bo_p = a.take<float>(Vp);
    size_t copies_end = a.used;
    float* Etab = reinterpret_cast<float*>(a.take<char>((size_t)Vp * Gp * e));   // forward operand type: fp16 in the bf16 mode (82 MB, L2-resident), fp32 in the fp32 mode
    assign(params, grads, m, v, sq, scal, WeT, W1xT, W1hT, W1h, W1x, W2xT, W2x, W2eT, W2e, W2hT, W2h, WoT, Wo, WembC, attrWT, be_p, b1_p, b2_p, bo_p, Etab,
           copies_begin, copies_end);
}

extern "C" int s2vt_create(const s2vt_config* cfg, s2vt_handle** out) {
    if (!cfg || !out) return S2VT_EINVAL;
    if (cfg->dim_image <= 0 || cfg->word_dim <= 0 || cfg->lstm_dim <= 0 || cfg->n_words <= 2 || cfg->n_video_steps <= 0 || cfg->n_caption_steps <= 0)
        return S2VT_EINVAL;
    if (cfg->n_words > ROW_THREADS * ROW_MAXV4 * 4 || cfg->n_words > 65534) return S2VT_EINVAL;
    if (cfg->precision != S2VT_PREC_BF16 && cfg->precision != S2VT_PREC_FP32) return S2VT_EINVAL;
    s2vt_handle* h = new s2vt_handle();
    h->cfg = *cfg;
    if (!(h->cfg.dropout_keep > 0.f) || h->cfg.dropout_keep > 1.f) h->cfg.dropout_keep = 1.f;
    h->D = cfg->dim_image; h->E = cfg->word_dim; h->H = cfg->lstm_dim; h->V = cfg->n_words;
    h->Tv = cfg->n_video_steps; h->Tc = cfg->n_caption_steps; h->A = cfg->n_attributes > 0 ? cfg->n_attributes : 0;
    h->T = h->Tv + h->Tc;
    h->Dp = ru(h->D, S2VT_PAD); h->Ep = ru(h->E, S2VT_PAD); h->Hp = ru(h->H, S2VT_PAD); h->Vp = ru(h->V, S2VT_PAD);
    h->Gp = 4 * h->Hp; h->Ap = h->A ? ru(h->A, S2VT_PAD) : 0;
    h->esz = cfg->precision == S2VT_PREC_BF16 ? 2 : 4;
    h->P = 0;
    const int64_t E = h->E, H = h->H, V = h->V, D = h->D;
    h->iWemb = 0; add_var(h, "Wemb", {}, V, E);
    h->iWe = 1; add_var(h, "encode_image_W", {}, D, E);
    h->ibe = 2; add_var(h, "encode_image_b", {}, E, 0);
    h->iWo = 3; add_var(h, "embed_word_W", {}, H, V);
    h->ibo = 4; add_var(h, "embed_word_b", {}, V, 0);
    h->iW1 = 5; add_var(h, "s2vt/LSTM1/basic_lstm_cell/weights", {"s2vt/LSTM1/basic_lstm_cell/kernel", "s2vt/LSTM1/BasicLSTMCell/Linear/Matrix"}, E + H, 4 * H);
    h->ib1 = 6; add_var(h, "s2vt/LSTM1/basic_lstm_cell/biases", {"s2vt/LSTM1/basic_lstm_cell/bias", "s2vt/LSTM1/BasicLSTMCell/Linear/Bias"}, 4 * H, 0);
    h->iW2 = 7; add_var(h, "s2vt/LSTM2/basic_lstm_cell/weights", {"s2vt/LSTM2/basic_lstm_cell/kernel", "s2vt/LSTM2/BasicLSTMCell/Linear/Matrix"}, H + E + H, 4 * H);
    h->ib2 = 8; add_var(h, "s2vt/LSTM2/basic_lstm_cell/biases", {"s2vt/LSTM2/basic_lstm_cell/bias", "s2vt/LSTM2/BasicLSTMCell/Linear/Bias"}, 4 * H, 0);
    h->iAW = h->iAb = -1;
    if (h->A) {
        h->iAW = 9; add_var(h, "attr_W", {}, D, h->A);
        h->iAb = 10; add_var(h, "attr_b", {}, h->A, 0);
    }
    Arena a(nullptr, 0);
    layout_state(h, a, [](auto...) {});
    h->state_bytes = a.used;
    h->state = nullptr; h->ws = nullptr; h->ws_bytes = 0; h->bound = false; h->fresh = false;
    *out = h;
    return S2VT_OK;
}

// Debug: device buffer of >= 8 * 4001 uint64 (zeroed by the caller) receiving per-launch phase timestamps, or NULL.
extern "C" int s2vt_debug_probe(void* device_buffer) {
    unsigned long long* p = (unsigned long long*)device_buffer;
    return cudaMemcpyToSymbol(tc::g_probe, &p, sizeof p) == cudaSuccess ? 0 : S2VT_ECUDA;
}
// Debug / tuning: which independent pieces run on the internal side stream (bit 0 late refresh, 1 dWo, 2 LSTM1 backward);
// bit 3 keeps beam search on its un-fused step (materialised logits, separate top-k / update / gather launches).
extern "C" int s2vt_set_overlap(s2vt_handle* h, int mask) {
    if (!h) return S2VT_EINVAL;
    h->overlap = mask;
    return 0;
}
extern "C" int s2vt_set_reuse_frontend(s2vt_handle* h, int enable) {
    if (!h) return S2VT_EINVAL;
    h->reuse_front = enable != 0;
    h->front_valid = false;
    return 0;
}
static int peer_close(s2vt_handle* h);
extern "C" void s2vt_destroy(s2vt_handle* h) {
    if (h && h->side) { cudaStreamDestroy(h->side); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); cudaEventDestroy(h->ev_refresh); cudaEventDestroy(h->ev_wo); cudaEventDestroy(h->ev_seg[0]); cudaEventDestroy(h->ev_seg[1]); cudaEventDestroy(h->ev_gate); }
    if (h) { peer_close(h); if (h->peer_comm) cudaFree(h->peer_comm); }
    if (h && h->tc_cache) delete static_cast<tc::MapCache*>(h->tc_cache);
    delete h;
}
extern "C" const char* s2vt_last_error(const s2vt_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" size_t s2vt_num_params(const s2vt_handle* h) { return h->P; }
extern "C" size_t s2vt_state_bytes(const s2vt_handle* h) { return h->state_bytes; }
extern "C" float* s2vt_params(const s2vt_handle* h) { return h->params; }
extern "C" float* s2vt_grads(const s2vt_handle* h) { return h->grads; }
extern "C" float* s2vt_adam_m(const s2vt_handle* h) { return h->adam_m; }
extern "C" float* s2vt_adam_v(const s2vt_handle* h) { return h->adam_v; }
extern "C" int s2vt_num_variables(const s2vt_handle* h) { return (int)h->vars.size(); }
extern "C" int s2vt_variable_info(const s2vt_handle* h, int index, const char** tf_name, int64_t* offset, int64_t shape[2], int* ndim) {
    if (index < 0 || index >= (int)h->vars.size()) return S2VT_EINVAL;
    const Var& v = h->vars[index];
    if (tf_name) *tf_name = v.name.c_str();
    if (offset) *offset = (int64_t)v.off;
    if (shape) { shape[0] = v.rows; shape[1] = v.cols; }
    if (ndim) *ndim = v.cols ? 2 : 1;
    return S2VT_OK;
}

extern "C" int s2vt_bind(s2vt_handle* h, void* state, size_t state_bytes, void* workspace, size_t workspace_bytes) {
    if (!h || !state) return S2VT_EINVAL;
    if (state_bytes < h->state_bytes) return h->fail(S2VT_ENOSPACE, "state block %zu < %zu bytes", state_bytes, h->state_bytes);
    if (((uintptr_t)state & 255) || ((uintptr_t)workspace & 255)) return h->fail(S2VT_EINVAL, "blocks must be 256-byte aligned");
    Arena a(state, state_bytes);
    layout_state(h, a, [&](float* params, float* grads, float* m, float* v, double* sq, float* scal, char* WeT, char* W1xT, char* W1hT, char* W1h, char* W1x,
                           char* W2xT, char* W2x, char* W2eT, char* W2e, char* W2hT, char* W2h, char* WoT, char* Wo, char* WembC, char* attrWT, float* be_p,
                           float* b1_p, float* b2_p, float* bo_p, float* Etab, size_t, size_t) {
        h->params = params; h->grads = grads; h->adam_m = m; h->adam_v = v; h->sq = sq; h->scal = scal; h->gbar = reinterpret_cast<unsigned*>(scal + 64);
        h->WeT = WeT; h->W1xT = W1xT; h->W1hT = W1hT; h->W1h = W1h; h->W1x = W1x; h->W2xT = W2xT; h->W2x = W2x; h->W2eT = W2eT; h->W2e = W2e;
        h->W2hT = W2hT; h->W2h = W2h; h->WoT = WoT; h->Wo = Wo; h->WembC = WembC; h->attrWT = attrWT;
        h->be_p = be_p; h->b1_p = b1_p; h->b2_p = b2_p; h->bo_p = bo_p; h->Etab = Etab;
    });
    h->state = (char*)state;
    h->ws = (char*)workspace; h->ws_bytes = workspace_bytes;
    h->bound = true; h->fresh = false;
    return S2VT_OK;
}

extern "C" int s2vt_load_param(s2vt_handle* h, const char* tf_name, const float* src_host, const int64_t* shape, int ndim, s2vt_stream st) {
    if (!h || !h->bound) return S2VT_ESTATE;
    for (const Var& v : h->vars) {
        bool match = v.name == tf_name;
        for (const auto& al : v.aliases) match |= (al == tf_name);
        if (!match) continue;
        int vnd = v.cols ? 2 : 1;
        if (ndim != vnd || shape[0] != v.rows || (vnd == 2 && shape[1] != v.cols)) return h->fail(S2VT_ESHAPE, "shape mismatch for %s", tf_name);
        CUDA_TRY(h, cudaMemcpyAsync(h->params + v.off, src_host, v.count() * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)st));
        h->fresh = false;
        return S2VT_OK;
    }
    return h->fail(S2VT_ENOTFOUND, "no variable named %s", tf_name);
}

// =================================================================================================================
// refresh: fp32 master -> compute copies
// =================================================================================================================
template <typename T>
static int pack(s2vt_handle* h, cudaStream_t st, const float* src, int lds, int R, int C, void* dst, int ldd, int gate_h, int transpose) {
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    pack_matrix_kernel<T><<<grid, block, 0, st>>>(src, lds, R, C, (T*)dst, ldd, gate_h, transpose);
    KCHECK(h);
    return 0;
}

static cudaEvent_t prof_event(s2vt_handle* h) {
    if (!h->prof_pool.empty()) { cudaEvent_t e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
// logical (un-padded) size of a padded dimension, for the algorithmic FLOP count of the roofline report
static double logical_dim(const s2vt_handle* h, int x) {
    if (x == h->Gp) return 4.0 * h->H;
    if (x == h->Hp) return h->H;
    if (x == h->Vp) return h->V;
    if (x == h->Ep) return h->E;
    if (x == h->Dp) return h->D;
    return x;
}
// Algorithmic bytes of the fused cell epilogues (logical sizes): what one per-step launch must move besides the
// GEMM operands.  Used only for the roofline report.
template <class Epi> struct EpiBytes { static double get(const s2vt_handle*, const typename Epi::Params&, int) { return 0.0; } };
template <typename T> struct EpiBytes<EpiLstmFwd<T>> {
    static double get(const s2vt_handle* h, const typename EpiLstmFwd<T>::Params& p, int M) {
        const double H = h->H, G = 4.0 * h->H;
        double b = M * H * 4 * 2 + M * H * sizeof(T);                       // c in, c out, h out
        if (p.add0) b += M * G * 4;
        if (p.add1) b += M * G * sizeof(T);                                 // Etab rows in the forward operand type
        if (p.gates_out) b += M * G * sizeof(T);
        if (p.hdrop_out) b += M * H * sizeof(T);
        return b;
    }
};
template <typename T, typename F> struct EpiBytes<EpiLstmBwd<T, F>> {
    static double get(const s2vt_handle* h, const typename EpiLstmBwd<T, F>::Params& p, int M) {
        const double H = h->H, G = 4.0 * h->H;
        return M * G * sizeof(T) + M * H * 4 * 4 + (p.a.dh_ext ? M * H * 4 : 0) + M * G * sizeof(T);   // gates, c_new/c_prev/dc in/out, dh_ext, dG out
    }
};
static int chain_begin(s2vt_handle* h, cudaStream_t st) {
    if (!h->prof) return 0;
    h->chain = s2vt_handle::ProfRec();
    h->chain.a = prof_event(h); h->chain.b = prof_event(h);
    h->chain.flops = h->chain.bytes = 0; h->chain.count = 0; h->chain.launches = 0; h->chain.cls = 1; h->chain.M = h->chain.N = h->chain.K = 0;
    h->chain_open = true;
    cudaEventRecord(h->chain.a, st);
    return 0;
}
static int chain_end(s2vt_handle* h, cudaStream_t st) {
    if (!h->prof || !h->chain_open) return 0;
    cudaEventRecord(h->chain.b, st);
    h->prof_recs.push_back(h->chain);
    h->chain_open = false;
    return 0;
}
// The vocabulary projection of a decode loop (fused pick / top-k epilogues) always follows that step's cell kernel, which never
// writes W_o: launched as a programmatic dependent, its prologue and the first W_o tiles overlap the cell kernel's tail.
template <class Epi> struct kPdlLogits { static constexpr bool value = false; };
template <typename T> struct kPdlLogits<EpiLogitsPick<T>> { static constexpr bool value = true; };
template <typename T> struct kPdlLogits<EpiLogitsTopK<T>> { static constexpr bool value = true; };
template <class Epi> struct IsCellBwd { static constexpr bool value = false; };
template <typename T, typename F> struct IsCellBwd<EpiLstmBwd<T, F>> { static constexpr bool value = true; };
// T: type of BOTH operands (forward products: Fwd<T>::type x Fwd<T>::type, backward products: T x T; the mixed weight-gradient
// products go through wgrad below).  16-bit types take the tcgen05 path, the operand format travels in the `fmt` word.
template <typename T, class Cfg, class Epi>
static int gemm(s2vt_handle* h, cudaStream_t st, const void* A, int lda, const void* B, int ldb, int M, int N, int K, const typename Epi::Params& ep,
                int logical_k = 0, int logical_m = 0) {
    constexpr uint32_t fmt = tc::FmtOf<T>::A | tc::FmtOf<T>::B;
    s2vt_handle::ProfRec rec;
    const bool in_chain = h->prof && h->chain_open && Cfg::BM < 128;
    const bool bracket = h->prof && !in_chain;
    if (h->prof) {
        const double lm = logical_m ? (double)logical_m : logical_dim(h, M), ln = logical_dim(h, N), lk = logical_k ? (double)logical_k : logical_dim(h, K);
        rec.flops = 2.0 * lm * ln * lk;
        rec.bytes = (lm * lk + ln * lk) * sizeof(T) + EpiBytes<Epi>::get(h, ep, M);   // operands once + epilogue traffic
        rec.cls = Cfg::BM >= 128 ? 0 : 1;
        rec.M = M; rec.N = N; rec.K = K; rec.count = 1; rec.launches = 1;
        if (in_chain) {
            h->chain.flops += rec.flops; h->chain.bytes += rec.bytes; h->chain.count += 1; h->chain.launches += 1;
            h->chain.M = M; h->chain.N = N; h->chain.K = K;
        } else {
            rec.a = prof_event(h); rec.b = prof_event(h);
            cudaEventRecord(rec.a, st);
        }
    }
    h->launches++;
    bool done = false;
    if constexpr (sizeof(T) == 2) {
        if (h->cfg.gemm_backend != S2VT_GEMM_MMA_SYNC) {   // tcgen05 + TMA + TMEM path (default for the 16-bit operand types)
            if (!h->tc_cache) h->tc_cache = new tc::MapCache();
            tc::MapCache& mc = *static_cast<tc::MapCache*>(h->tc_cache);
            // per-step GEMMs are launched as programmatic dependents: their prologue and weight prefetch overlap the tail of the
            // preceding kernel (which never writes weights: only s2vt_refresh does, and a batched GEMM always follows it)
            if (Cfg::BM >= 128) {
                // batched GEMMs are bound by L2 -> SM bandwidth, so use the widest tile (128 x 256: 85 flop per byte fetched vs
                // 64 for 128 x 128).  gemm_backend 3 / 4 select the 128-wide tile without / with 2x2 TMA multicast (measured
                // slower on B200: at cluster sizes <= 4 multicast does not reduce L2 traffic).
                // large products: 256 x 256 tiles on CTA pairs (cta_group::2, 131 flop per fetched byte); gemm_backend 14 keeps them on single CTAs
                bool paired = false;
                if constexpr (!Epi::kDirect && !kPdlLogits<Epi>::value) {
                    // isolated timings (scripts/gemm_shapes.py, profiles/r2_gemm_shapes.md): the pair tile wins where the single-CTA tile is fill-bound --
                    // long contractions (K >= 8192: 1375 vs 1178 TF/s) and narrow outputs with K >= 4096; wide short-K products are bound by
                    // their fp32 stores and the MMA rate, where the pair only adds its cluster hand-shakes
                    const bool pair_wins = K >= 8192 || (K >= 4096 && N <= 512);
                    if ((h->cfg.gemm_backend == S2VT_GEMM_AUTO || h->cfg.gemm_backend == S2VT_GEMM_TCGEN05 || h->cfg.gemm_backend == 15) && N % 256 == 0 && M >= 1024 &&
                        (pair_wins || h->cfg.gemm_backend == 15)) {
                        CUDA_TRY(h, (tc::launch_pair<Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false, fmt)));
                        paired = true;
                    }
                }
                // wide short-K products with fp32 output (G1x, G2x, logits): persistent tile loop, the stores of tile i under the mainloop of tile i+1
                if constexpr (std::is_same<Epi, EpiStore<T>>::value) {
                    if (!paired && (h->cfg.gemm_backend == S2VT_GEMM_AUTO || h->cfg.gemm_backend == S2VT_GEMM_TCGEN05 || h->cfg.gemm_backend == 15) && N % 256 == 0 &&
                        (long long)((M + 127) / 128) * (N / 256) >= 296 && (ep.outF || ep.outT) && !ep.accumulate && ep.M == M) {
                        CUDA_TRY(h, (tc::launch_persist<T>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep.outF, ep.outT, ep.ldo, ep.bias, false, fmt)));
                        paired = true;
                    }
                }
                if (paired) {}
                else if (h->cfg.gemm_backend == 4 && (N / 128) % 2 == 0) CUDA_TRY(h, (tc::launch<128, Epi, 1, 2, 2>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false, fmt)));
                else if (h->cfg.gemm_backend == 3 || N % 256 != 0) CUDA_TRY(h, (tc::launch<128, Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, false, fmt)));
                else CUDA_TRY(h, (tc::launch<256, Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, kPdlLogits<Epi>::value, fmt)));
            }
            else if constexpr (IsCellBwd<Epi>::value) {
                // cell backward: K = 4H is long and N = H gives few tiles -> split K over a cluster of 4 CTAs (DSMEM reduction)
                if ((K / tc::BK) % 4 == 0) {
                    if (M > 128 && h->cfg.gemm_backend != 7) CUDA_TRY(h, (tc::launch<128, Epi, 4>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
                    else if (M > 128) CUDA_TRY(h, (tc::launch<64, Epi, 4>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
                    else CUDA_TRY(h, (tc::launch<32, Epi, 4>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
                } else CUDA_TRY(h, (tc::launch<32, Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
            }
            else if (M > 128 && N > 1024) {
                // every column tile of a row block reads the same activations: fetch them once per cluster of 8 (TMA multicast)
                if (h->cfg.gemm_backend == 5 && (N / 64) % 8 == 0) CUDA_TRY(h, (tc::launch<64, Epi, 1, 8, 1>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
                else if (h->cfg.gemm_backend == 6 && (N / 64) % 4 == 0) CUDA_TRY(h, (tc::launch<64, Epi, 1, 4, 1>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
                else if (h->cfg.gemm_backend == 7) CUDA_TRY(h, (tc::launch<64, Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
                else CUDA_TRY(h, (tc::launch<128, Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));   // <= 148 CTAs, one per SM: balanced smem fill
            }
            else CUDA_TRY(h, (tc::launch<32, Epi>(mc, st, (const bf16*)A, lda, (const bf16*)B, ldb, M, N, K, ep, true, fmt)));
            done = true;
        }
    }
    if (!done) CUDA_TRY(h, (launch_gemm<T, Cfg, Epi>(st, (const T*)A, lda, (const T*)B, ldb, M, N, K, ep)));
    if (bracket) { cudaEventRecord(rec.b, st); h->prof_recs.push_back(rec); }
    return 0;
}

// Weight gradient  grad (+)= X^T . Y  over R rows (X [R, ldx]: Mf padded features, Y [R, ldy]: Nf padded features).
// tcgen05 path: MN-major operand descriptors read X and Y as they lie (no transposed copies, ragged R zero-filled by TMA).
// Other mainloops need K-major operands: X and Y are transposed into the scratch buffers first.
template <typename TI, typename TO> static int transpose(s2vt_handle* h, cudaStream_t st, const TI* src, int lds, int R, int C, TO* dst, int ldd, int rows_dst_padded);
#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)
// X: a forward activation (type F = Fwd<T>::type), Y: a gradient (type T).  tcgen05.mma kind::f16 wants ONE format for both operands
// (an instruction descriptor with a_format F16 and b_format BF16 raises an illegal-instruction fault on B200 -- measured), so in
// the bf16 mode the fp16 activations are rounded to bf16 on their way into the product (one streaming pass, ~0.4 GB per iteration).
template <typename T, typename F>
static int wgrad(s2vt_handle* h, cudaStream_t st, const F* X, int ldx, int Mf, const T* Y, int ldy, int Nf, int R, const EpiGradStore::Params& ep,
                 T* tA, T* tB, int logical_m) {
    if constexpr (sizeof(T) == 2) {
        constexpr uint32_t fmt = tc::FmtOf<T>::A | tc::FmtOf<T>::B;
        if (h->cfg.gemm_backend != S2VT_GEMM_MMA_SYNC && h->cfg.gemm_backend != 8) {
            if constexpr (!std::is_same<F, T>::value) {
                const size_t n = (size_t)R * ldx;
                convert_kernel<F, T><<<148 * 8, 256, 0, st>>>(X, n, tA); KCHECK(h);      // ldx is a multiple of 128: n % 8 == 0
                X = reinterpret_cast<const F*>(tA);
            }
            s2vt_handle::ProfRec rec;
            if (h->prof) {
                rec.a = prof_event(h); rec.b = prof_event(h);
                rec.flops = 2.0 * logical_m * logical_dim(h, Nf) * (double)R;
                rec.bytes = ((double)logical_m + logical_dim(h, Nf)) * R * sizeof(T);
                rec.cls = 0; rec.M = Mf; rec.N = Nf; rec.K = R; rec.count = 1; rec.launches = 1;
                cudaEventRecord(rec.a, st);
            }
            h->launches++;
            if (!h->tc_cache) h->tc_cache = new tc::MapCache();
            tc::MapCache& mc = *static_cast<tc::MapCache*>(h->tc_cache);
            if (Nf % 256 == 0 && Mf >= 256 && (h->cfg.gemm_backend == S2VT_GEMM_AUTO || h->cfg.gemm_backend == S2VT_GEMM_TCGEN05 || h->cfg.gemm_backend == 15)) {
                // few output tiles (e.g. dWe: 24, dW2[emb rows]: 64 on 148 SMs): split the long contraction over grid z, partial tiles are added atomically
                const int tiles = ((Mf + 255) / 256) * 2 * (Nf / 256), kblocks = (R + tc::BK - 1) / tc::BK;
                int ks = tiles >= 100 ? 1 : 148 / tiles;
                if (ks > kblocks / 8) ks = kblocks / 8 > 0 ? kblocks / 8 : 1;
                EpiGradStore::Params eps = ep;
                eps.atomic = ks > 1;
                CUDA_TRY(h, (tc::launch_pair<EpiGradStore, true>(mc, st, (const bf16*)X, ldx, (const bf16*)Y, ldy, Mf, Nf, R, eps, false, fmt, ks)));
            }
            else if (Nf % 256 == 0) CUDA_TRY(h, (tc::launch<256, EpiGradStore, 1, 1, 1, true>(mc, st, (const bf16*)X, ldx, (const bf16*)Y, ldy, Mf, Nf, R, ep, false, fmt)));
            else CUDA_TRY(h, (tc::launch<128, EpiGradStore, 1, 1, 1, true>(mc, st, (const bf16*)X, ldx, (const bf16*)Y, ldy, Mf, Nf, R, ep, false, fmt)));
            if (h->prof) { cudaEventRecord(rec.b, st); h->prof_recs.push_back(rec); }
            return 0;
        }
    }
    const int Rp = ru(R, S2VT_PAD);
    TRY((transpose<F, T>(h, st, X, ldx, R, Mf, tA, Rp, Mf)));     // the K-major mainloops multiply like types: X is converted on the way
    TRY((transpose<T, T>(h, st, Y, ldy, R, Nf, tB, Rp, Nf)));
    return gemm<T, CfgBig, EpiGradStore>(h, st, tA, Rp, tB, Rp, Mf, Nf, Rp, ep, R, logical_m);
}
template <typename T>
static int bias_grad(s2vt_handle* h, cudaStream_t st, const T* Y, int ldy, int Cpad, int R, int ncols, int gate_h, float* grad) {
    dim3 grid(Cpad / 64, R >= 16384 ? 32 : (R >= 4096 ? 16 : (R >= 256 ? 4 : 1))), block(32, 8);
    colsum_grad_kernel<T><<<grid, block, 0, st>>>(Y, ldy, R, ncols, gate_h, grad);
    KCHECK(h);
    return 0;
}

// A chain of dependent per-step GEMMs (one LSTM layer walked over time).  bf16 / tcgen05: ONE persistent launch with grid
// barriers between the steps (gemm_tcgen05_chain.cuh); otherwise (fp32, mma.sync checker, grid too large to be co-resident,
// gemm_backend 9) one launch per step.
template <typename T, class Epi>
struct StepChain {
    std::vector<typename Epi::Params> eps;
    const T* A; int lda, a_total_rows, a_row0, a_row_stride;   // step s reads rows [a_row0 + s * a_row_stride, + M) of A
    const T* B; int ldb, M, N, K;
    float* ws2_scratch = nullptr; unsigned* ws2_flags = nullptr; int ws2_flags_cap = 0;   // backward chains > 128 rows: partial-tile scratch / counters (gemm_tcgen05_ws2.cuh)
    int gate_ncta = 0;    // out: CTAs of the persistent launch when its progress can be watched on h->gbar (count >= s * gate_ncta <=> steps 0..s-1 are complete), else 0
};
template <typename T, class Epi>
static int run_chain(s2vt_handle* h, cudaStream_t st, StepChain<T, Epi>& c, void* dev_params) {
    const int n = (int)c.eps.size();
    if (n == 0) return 0;
    chain_begin(h, st);
    bool done = false;
    if constexpr (sizeof(T) == 2) {
        constexpr uint32_t fmt = tc::FmtOf<T>::A | tc::FmtOf<T>::B;
        if (h->cfg.gemm_backend != S2VT_GEMM_MMA_SYNC && h->cfg.gemm_backend != 9 && n >= 2) {
            if (!h->tc_cache) h->tc_cache = new tc::MapCache();
            tc::MapCache& mc = *static_cast<tc::MapCache*>(h->tc_cache);
            typedef typename Epi::Params P;
            CUDA_TRY(h, cudaMemcpyAsync(dev_params, c.eps.data(), (size_t)n * sizeof(P), cudaMemcpyHostToDevice, st));
            unsigned* gbar = h->gbar + (st == h->side ? 16 : 0);
            const P* dp = (const P*)dev_params;
            cudaError_t e;
            constexpr bool bwd = IsCellBwd<Epi>::value;
            if constexpr (bwd) {
                if ((c.K / tc::BK) % 4 != 0) e = cudaErrorLaunchOutOfResources;
                else if (c.M > 128) {
                    // > 128 rows: the ring chain with its 4-CTA DSMEM exchange; alternative: weights-stationary K-slice slabs, two pipelined halves, partial tiles
                    // reduced through L2 (gemm_tcgen05_ws2.cuh).
                    e = cudaErrorLaunchOutOfResources;
                    // Measured at the bench shape (round 2): 16.3 us per step against 15.3 us for the ring chain inside the iteration -- the L2 round trip of the
                    // partial tiles and the 132-CTA grid (which leaves the concurrent dW_o GEMM 16 SMs) cost more than the resident weights save: opt-in, gemm_backend 16.
                    if (h->cfg.gemm_backend == 16 && c.ws2_scratch)
                        e = tc::launch_ws2_bwd_chain<Epi>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n,
                                                          c.ws2_flags, c.ws2_flags_cap, c.ws2_scratch, true, fmt);
                    if (e == cudaErrorLaunchOutOfResources) {
                        (void)cudaGetLastError();
                        e = tc::launch_chain<128, Epi, 4>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt);
                        if (e == cudaSuccess) c.gate_ncta = (c.N / 128) * ((c.M + tc::BM - 1) / tc::BM) * 4;
                    }
                }
                else {
                    e = cudaErrorLaunchOutOfResources;
                    if (c.M <= 64 && h->cfg.gemm_backend != 10) e = tc::launch_chain<32, Epi, 4, true>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt);
                    if (e == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); e = tc::launch_chain<32, Epi, 4>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt); }
                }
            } else {
                if (c.M > 128) {
                    // > 128 rows: weights-stationary slabs + two software-pipelined halves per row group (gemm_tcgen05_ws2.cuh);
                    // the plain ring chain is the fallback (shape does not fit / gemm_backend 13)
                    e = cudaErrorLaunchOutOfResources;
                    static const int ws2_bn64 = getenv("S2VT_WS2_BN64") ? atoi(getenv("S2VT_WS2_BN64")) : 0;      // experiment: 64-column slabs, 2 row groups, 12-stage ring
                    if (ws2_bn64 && c.M <= 256 && h->cfg.gemm_backend != 13 && h->cfg.gemm_backend != 10)
                        e = tc::launch_ws2_chain<Epi, 64>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n,
                                                          gbar + 32, true, fmt);
                    else if (h->cfg.gemm_backend != 13 && h->cfg.gemm_backend != 10)
                        e = tc::launch_ws2_chain<Epi>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n,
                                                      gbar + 32, true, fmt);
                    if (e == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); e = tc::launch_chain<128, Epi, 1>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt); }
                }
                else {
                    e = cudaErrorLaunchOutOfResources;   // weights-stationary variant first (rows <= 64, K <= 16 blocks)
                    // weights stationary + the activation rows multicast over clusters of 8 column tiles (gemm_backend 11 / 12: clusters of 4 / none)
                    if (c.M <= 64 && h->cfg.gemm_backend != 10 && h->cfg.gemm_backend != 11 && h->cfg.gemm_backend != 12) e = tc::launch_chain<32, Epi, 1, true, 8>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt);
                    if (e == cudaErrorLaunchOutOfResources && c.M <= 64 && h->cfg.gemm_backend == 11) { (void)cudaGetLastError(); e = tc::launch_chain<32, Epi, 1, true, 4>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt); }
                    if (e == cudaErrorLaunchOutOfResources && c.M <= 64 && h->cfg.gemm_backend != 10) { (void)cudaGetLastError(); e = tc::launch_chain<32, Epi, 1, true>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt); }
                    if (e == cudaErrorLaunchOutOfResources) { (void)cudaGetLastError(); e = tc::launch_chain<32, Epi, 1>(mc, st, (const bf16*)c.A, c.lda, c.a_total_rows, c.a_row0, c.a_row_stride, (const bf16*)c.B, c.ldb, c.M, c.N, c.K, dp, n, gbar, true, fmt); }
                }
            }
            if (e == cudaSuccess) {
                done = true;
                h->launches++;
                if (h->prof && h->chain_open) {
                    for (int s2 = 0; s2 < n; ++s2) {
                        const double lm = logical_dim(h, c.M), ln = logical_dim(h, c.N), lk = logical_dim(h, c.K);
                        h->chain.flops += 2.0 * lm * ln * lk;
                        h->chain.bytes += (lm * lk + ln * lk) * sizeof(T) + EpiBytes<Epi>::get(h, c.eps[s2], c.M);
                    }
                    h->chain.count += n; h->chain.launches += 1; h->chain.M = c.M; h->chain.N = c.N; h->chain.K = c.K;
                }
            } else if (e != cudaErrorLaunchOutOfResources) {
                return h->fail(S2VT_ECUDA, "persistent chain launch failed: %s", cudaGetErrorString(e));
            } else {
                (void)cudaGetLastError();
            }
        }
    }
    if (!done) {
        for (int s2 = 0; s2 < n; ++s2)
            TRY((gemm<T, CfgStep, Epi>(h, st, c.A + (size_t)(c.a_row0 + s2 * c.a_row_stride) * c.lda, c.lda, c.B, c.ldb, c.M, c.N, c.K, c.eps[s2])));
    }
    chain_end(h, st);
    return 0;
}

static int ensure_side(s2vt_handle* h) {
    if (h->side) return 0;
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_refresh, cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_wo, cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_seg[0], cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_seg[1], cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_gate, cudaEventDisableTiming));
    return 0;
}
// The "late" half of a refresh (everything but the frame projection / LSTM1 forward weights) runs on the side stream;
// consumers on the caller's stream wait for it here.  A no-op when no refresh is in flight.
static int wait_late_weights(s2vt_handle* h, cudaStream_t st) {
    if (h->ev_refresh) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_refresh, 0));
    return 0;
}

template <typename T>
static int refresh_impl(s2vt_handle* h, cudaStream_t st) {
    const int E = h->E, H = h->H, V = h->V, D = h->D, Dp = h->Dp, Ep = h->Ep, Hp = h->Hp, Vp = h->Vp, Gp = h->Gp;
    // zero all compute copies (pads must be zero)
    size_t begin = 0, end = 0;
    {
        Arena a(nullptr, 0);
        layout_state(h, a, [&](float*, float*, float*, float*, double*, float*, char*, char*, char*, char*, char*, char*, char*, char*, char*, char*, char*,
                               char*, char*, char*, char*, float*, float*, float*, float*, float*, size_t b, size_t e) { begin = b; end = e; });
    }
    if (!h->copies_zeroed) {   // the pads are never written afterwards
        CUDA_TRY(h, cudaMemsetAsync(h->state + begin, 0, end - begin, st));
        h->copies_zeroed = true;
    }
    typedef typename Fwd<T>::type F;      // K-major copies feed the forward products (fp16 in the bf16 mode), TF-layout copies the backward ones
    const float* W1 = h->P_(h->iW1); const float* W2 = h->P_(h->iW2);
    const int G = 4 * H;
    TRY(ensure_side(h));
    cudaStream_t s2 = (h->overlap & 1) ? h->side : st;
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));                 // parameters (and the zeroing) are final on `st` here
    CUDA_TRY(h, cudaStreamWaitEvent(s2, h->ev_fork, 0));
    // early half, caller's stream: what the next call needs first (frame projection and LSTM1 forward)
    TRY(pack<F>(h, st, h->P_(h->iWe), E, D, E, h->WeT, Dp, 0, 1));                         // WeT [Ep, Dp]
    TRY(pack<F>(h, st, W1, G, E, G, h->W1xT, Ep, H, 1));                                   // W1xT [Gp, Ep]
    TRY(pack<F>(h, st, W1 + (size_t)E * G, G, H, G, h->W1hT, Hp, H, 1));                   // W1hT [Gp, Hp]
    pack_vector_kernel<<<(E + 255) / 256, 256, 0, st>>>(h->P_(h->ibe), E, h->be_p, 0); KCHECK(h);
    pack_vector_kernel<<<(G + 255) / 256, 256, 0, st>>>(h->P_(h->ib1), G, h->b1_p, H); KCHECK(h);
    // late half, side stream: overlaps the next call's frame projection and LSTM1 chain (see wait_late_weights)
    TRY(pack<T>(h, s2, W1, G, E, G, h->W1x, Gp, H, 0));                                    // W1x  [Ep, Gp]
    TRY(pack<T>(h, s2, W1 + (size_t)E * G, G, H, G, h->W1h, Gp, H, 0));                    // W1h  [Hp, Gp]
    TRY(pack<F>(h, s2, W2, G, H, G, h->W2xT, Hp, H, 1));                                   // rows [0,H): out1
    TRY(pack<T>(h, s2, W2, G, H, G, h->W2x, Gp, H, 0));
    TRY(pack<F>(h, s2, W2 + (size_t)H * G, G, E, G, h->W2eT, Ep, H, 1));                   // rows [H,H+E): word embedding
    TRY(pack<T>(h, s2, W2 + (size_t)H * G, G, E, G, h->W2e, Gp, H, 0));
    TRY(pack<F>(h, s2, W2 + (size_t)(H + E) * G, G, H, G, h->W2hT, Hp, H, 1));             // rows [H+E, 2H+E): h2
    TRY(pack<T>(h, s2, W2 + (size_t)(H + E) * G, G, H, G, h->W2h, Gp, H, 0));
    TRY(pack<F>(h, s2, h->P_(h->iWo), V, H, V, h->WoT, Hp, 0, 1));                         // WoT [Vp, Hp]
    TRY(pack<T>(h, s2, h->P_(h->iWo), V, H, V, h->Wo, Vp, 0, 0));                          // Wo  [Hp, Vp]
    TRY(pack<F>(h, s2, h->P_(h->iWemb), E, V, E, h->WembC, Ep, 0, 0));                     // Wemb [Vp, Ep]
    if (h->A) TRY(pack<F>(h, s2, h->P_(h->iAW), h->A, D, h->A, h->attrWT, Dp, 0, 1));      // attrWT [Ap, Dp]
    pack_vector_kernel<<<(G + 255) / 256, 256, 0, s2>>>(h->P_(h->ib2), G, h->b2_p, H); KCHECK(h);
    pack_vector_kernel<<<(V + 255) / 256, 256, 0, s2>>>(h->P_(h->ibo), V, h->bo_p, 0); KCHECK(h);
    // Etab[v, :] = Wemb[v, :] . W2[emb rows]  (packed gate order) -- the word-embedding contribution to LSTM2's gates
    typename EpiStore<F>::Params ep = {nullptr, nullptr, Gp, nullptr, Vp, 0};
    if (std::is_same<F, float>::value) ep.outF = h->Etab; else ep.outT = reinterpret_cast<F*>(h->Etab);
    TRY((gemm<F, CfgBig, EpiStore<F>>(h, s2, h->WembC, Ep, h->W2eT, Ep, Vp, Gp, Ep, ep)));
    CUDA_TRY(h, cudaEventRecord(h->ev_refresh, s2));
    h->fresh = true;
    h->front_valid = false;
    return 0;
}

extern "C" int s2vt_refresh(s2vt_handle* h, s2vt_stream st) {
    if (!h || !h->bound) return S2VT_ESTATE;
    return h->cfg.precision == S2VT_PREC_BF16 ? refresh_impl<bf16>(h, (cudaStream_t)st) : refresh_impl<float>(h, (cudaStream_t)st);
}

// =================================================================================================================
// forward plans
// =================================================================================================================
template <typename T>
struct Front {   // per-video part: frame projection + LSTM1 over all T steps (forward values: type F)
    typedef typename Fwd<T>::type F;
    F* Xc; F* img; float* G1x; F* h1_all; float* c1_all; F* gates1; void* chain;
};
template <typename T>
static void plan_front(const s2vt_handle* h, Arena& a, int B, bool train, Front<T>& f) {
    typedef typename Fwd<T>::type F;
    f.Xc = a.take<F>((size_t)h->Tv * B * h->Dp);
    f.img = a.take<F>((size_t)h->Tv * B * h->Ep);
    f.G1x = a.take<float>((size_t)h->Tv * B * h->Gp);
    f.h1_all = a.take<F>((size_t)(h->T + 1) * B * h->Hp);
    f.c1_all = a.take<float>((size_t)(h->T + 1) * B * h->Hp);
    f.gates1 = train ? a.take<F>((size_t)h->T * B * h->Gp) : nullptr;
    f.chain = a.take<char>((size_t)(h->T + 1) * 512);   // per-step epilogue parameters of the persistent chains (<= 512 B each)
}

template <typename T>
static int run_front(s2vt_handle* h, cudaStream_t st, const float* video, int B, Front<T>& f) {
    typedef typename Fwd<T>::type F;
    const int Tv = h->Tv, T_ = h->T, Dp = h->Dp, Ep = h->Ep, Hp = h->Hp, Gp = h->Gp;
    convert_video_kernel<F><<<Tv * B, 256, 0, st>>>(video, nullptr, B, Tv, h->D, Dp, f.Xc); KCHECK(h);
    {   // img = Xc . We + be   (:107-111 tf.nn.xw_plus_b)
        typename EpiStore<F>::Params ep = {nullptr, f.img, Ep, h->be_p, Tv * B, 0};
        TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, f.Xc, Dp, h->WeT, Dp, Tv * B, Ep, Dp, ep)));
    }
    {   // G1x = img . W1[x rows]  (the input half of LSTM1's concat([x, h]) W, all frames at once)
        typename EpiStore<F>::Params ep = {f.G1x, nullptr, Gp, nullptr, Tv * B, 0};
        TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, f.img, Ep, h->W1xT, Ep, Tv * B, Gp, Ep, ep)));
    }
    CUDA_TRY(h, cudaMemsetAsync(f.h1_all, 0, (size_t)B * Hp * sizeof(F), st));
    CUDA_TRY(h, cudaMemsetAsync(f.c1_all, 0, (size_t)B * Hp * sizeof(float), st));
    StepChain<F, EpiLstmFwd<F>> ch;
    ch.A = f.h1_all; ch.lda = Hp; ch.a_total_rows = (T_ + 1) * B; ch.a_row0 = 0; ch.a_row_stride = B;
    ch.B = (const F*)h->W1hT; ch.ldb = Hp; ch.M = B; ch.N = Gp; ch.K = Hp;
    for (int t = 0; t < T_; ++t) {   // :128-129 / :148-149 LSTM1; decoder steps get `padding` -> no input term (Q8)
        typename EpiLstmFwd<F>::Params ep;
        memset(&ep, 0, sizeof ep);
        ep.M = B; ep.Hp = Hp; ep.bias = h->b1_p;
        ep.add0 = t < Tv ? f.G1x + (size_t)t * B * Gp : nullptr;
        ep.c_prev = f.c1_all + (size_t)t * B * Hp; ep.c_out = f.c1_all + (size_t)(t + 1) * B * Hp;
        ep.h_out = f.h1_all + (size_t)(t + 1) * B * Hp;
        ep.gates_out = f.gates1 ? f.gates1 + (size_t)t * B * Gp : nullptr;
        ep.keep = 1.f;
        ch.eps.push_back(ep);
    }
    TRY((run_chain<F, EpiLstmFwd<F>>(h, st, ch, f.chain)));
    return 0;
}

// ---- rollout (greedy + K samples), also the front half of beam search --------------------------------------------
template <typename T>
struct Roll {
    typedef typename Fwd<T>::type F;
    Front<T> f; float* G2x; F* h2enc; float* c2e[2]; F* h2r[2]; float* c2r[2]; float* logits; int* tok[2]; int* ids; void* chain;
    F* h2_final;   // encoder LSTM2 output state (rows of h2enc after the last frame)
    float* pick_val; int* pick_idx; int pick_ld;
};
template <typename T>
static void plan_roll(const s2vt_handle* h, Arena& a, int B, int R, Roll<T>& r) {
    typedef typename Fwd<T>::type F;
    plan_front<T>(h, a, B, true, r.f);   // same layout as the training plan so the LSTM1 forward can be shared
    r.G2x = a.take<float>((size_t)h->T * B * h->Gp);
    r.h2enc = a.take<F>((size_t)(h->Tv + 1) * B * h->Hp);   // LSTM2 state after every frame (one buffer: the persistent chain strides through it)
    r.h2_final = r.h2enc ? r.h2enc + (size_t)h->Tv * B * h->Hp : nullptr;
    for (int i = 0; i < 2; ++i) r.c2e[i] = a.take<float>((size_t)B * h->Hp);
    r.chain = a.take<char>((size_t)(h->T + 1) * 512);
    for (int i = 0; i < 2; ++i) { r.h2r[i] = a.take<F>((size_t)R * h->Hp); r.c2r[i] = a.take<float>((size_t)R * h->Hp); }
    r.logits = a.take<float>((size_t)R * h->Vp);
    r.tok[0] = a.take<int>(R); r.tok[1] = a.take<int>(R);
    r.ids = a.take<int>((size_t)R * h->Tc);
    r.pick_ld = h->Vp / 128;
    r.pick_val = a.take<float>((size_t)R * r.pick_ld); r.pick_idx = a.take<int>((size_t)R * r.pick_ld);
}

// frames -> (LSTM1 all steps, G2x all steps, LSTM2 encoder steps).  Leaves the encoder state in h2_final, c2e[Tv&1].
template <typename T>
static int run_encoder(s2vt_handle* h, cudaStream_t st, const float* video, int B, Roll<T>& r) {
    typedef typename Fwd<T>::type F;
    const int Tv = h->Tv, T_ = h->T, Hp = h->Hp, Gp = h->Gp;
    h->front_valid = false;
    TRY(run_front<T>(h, st, video, B, r.f));
    h->front_valid = true; h->front_B = B; h->front_video = video;
    TRY(wait_late_weights(h, st));
    {   // G2x = h1 . W2[out1 rows] for every step (bare cells: no dropout in the samplers, Q2)
        typename EpiStore<F>::Params ep = {r.G2x, nullptr, Gp, nullptr, T_ * B, 0};
        TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, r.f.h1_all + (size_t)B * Hp, Hp, h->W2xT, Hp, T_ * B, Gp, Hp, ep)));
    }
    CUDA_TRY(h, cudaMemsetAsync(r.h2enc, 0, (size_t)B * Hp * sizeof(F), st));
    CUDA_TRY(h, cudaMemsetAsync(r.c2e[0], 0, (size_t)B * Hp * sizeof(float), st));
    StepChain<F, EpiLstmFwd<F>> ch;
    ch.A = r.h2enc; ch.lda = Hp; ch.a_total_rows = (Tv + 1) * B; ch.a_row0 = 0; ch.a_row_stride = B;
    ch.B = (const F*)h->W2hT; ch.ldb = Hp; ch.M = B; ch.N = Gp; ch.K = Hp;
    for (int t = 0; t < Tv; ++t) {   // :131-132 LSTM2 on concat([output1, padding]) -> the embedding rows see zeros (Q8)
        typename EpiLstmFwd<F>::Params ep;
        memset(&ep, 0, sizeof ep);
        ep.M = B; ep.Hp = Hp; ep.bias = h->b2_p; ep.add0 = r.G2x + (size_t)t * B * Gp;
        ep.c_prev = r.c2e[t & 1]; ep.c_out = r.c2e[(t + 1) & 1]; ep.h_out = r.h2enc + (size_t)(t + 1) * B * Hp; ep.keep = 1.f;
        ch.eps.push_back(ep);
    }
    TRY((run_chain<F, EpiLstmFwd<F>>(h, st, ch, r.chain)));
    return 0;
}

// tile width the vocabulary-projection GEMM will use (decides how many candidates per row the fused pick / top-k epilogues produce)
template <typename T>
static int logits_tile_bn(const s2vt_handle* h) {
    const bool tc_path = std::is_same<T, bf16>::value && h->cfg.gemm_backend != S2VT_GEMM_MMA_SYNC;
    return tc_path && h->cfg.gemm_backend != 3 && h->cfg.gemm_backend != 4 && h->Vp % 256 == 0 ? 256 : 128;
}

template <typename T>
static int rollout_impl(s2vt_handle* h, cudaStream_t st, const float* video, int B, int K, uint64_t seed, uint32_t row_base, int32_t* sampled_out,
                        int32_t* greedy_out) {
    typedef typename Fwd<T>::type F;
    const int Tv = h->Tv, Tc = h->Tc, Hp = h->Hp, Gp = h->Gp, Vp = h->Vp;
    const bool want_greedy = greedy_out != nullptr;
    const int R = (K + (want_greedy ? 1 : 0)) * B;
    if (R <= 0) return h->fail(S2VT_EINVAL, "nothing to decode");
    Arena a(h->ws, h->ws_bytes);
    Roll<T> r;
    plan_roll<T>(h, a, B, R, r);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small: need %zu bytes", a.used);
    TRY(run_encoder<T>(h, st, video, B, r));
    tile_rows_kernel<F><<<R, 256, 0, st>>>(r.h2_final, B, R, Hp, r.h2r[0]); KCHECK(h);
    tile_rows_kernel<float><<<R, 256, 0, st>>>(r.c2e[Tv & 1], B, R, Hp, r.c2r[0]); KCHECK(h);
    fill_int_kernel<<<(R + 255) / 256, 256, 0, st>>>(r.tok[0], R, 1); KCHECK(h);   // <bos> = 1 (:321-323)
    const int logits_bn = logits_tile_bn<T>(h);
    const int nt = Vp / logits_bn;
    std::vector<typename EpiLstmFwd<F>::Params> cells(Tc);
    std::vector<typename EpiLogitsPick<T>::Params> picks(Tc);
    for (int i = 0; i < Tc; ++i) {
        const int t = Tv + i;
        typename EpiLstmFwd<F>::Params& ep = cells[i];
        memset(&ep, 0, sizeof ep);
        ep.M = R; ep.Hp = Hp; ep.bias = h->b2_p; ep.add0 = r.G2x + (size_t)t * B * Gp; ep.add0_mod = B;
        ep.add1 = reinterpret_cast<const F*>(h->Etab); ep.tok = r.tok[0];                     // step 0: <bos>; later steps resolve the previous step's candidates
        if (i > 0) { ep.pick_val = r.pick_val; ep.pick_idx = r.pick_idx; ep.pick_ld = r.pick_ld; ep.pick_nt = nt; ep.ids_out = r.ids; ep.ids_ld = Tc; ep.ids_col = i - 1; }
        ep.c_prev = r.c2r[i & 1]; ep.c_out = r.c2r[(i + 1) & 1]; ep.h_out = r.h2r[(i + 1) & 1]; ep.keep = 1.f;
        // :332-336 logit_words -> log_softmax -> tf.multinomial / :386-387 argmax, fused: the logits stay on chip
        typename EpiLogitsPick<T>::Params el = {R, h->V, h->bo_p, K * B, seed, (uint32_t)i, row_base, r.pick_val, r.pick_idx, r.pick_ld};
        picks[i] = el;
    }
    for (int i = 0; i < Tc; ++i) {
        TRY((gemm<F, CfgStep, EpiLstmFwd<F>>(h, st, r.h2r[i & 1], Hp, h->W2hT, Hp, R, Gp, Hp, cells[i])));
        TRY((gemm<F, CfgBig, EpiLogitsPick<T>>(h, st, r.h2r[(i + 1) & 1], Hp, h->WoT, Hp, R, Vp, Hp, picks[i])));
    }
    resolve_picks_kernel<<<(R + 127) / 128, 128, 0, st>>>(r.pick_val, r.pick_idx, r.pick_ld, nt, R, r.ids, Tc, Tc - 1); KCHECK(h);
    if (K > 0 && sampled_out)
        CUDA_TRY(h, cudaMemcpyAsync(sampled_out, r.ids, (size_t)K * B * Tc * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (want_greedy)
        CUDA_TRY(h, cudaMemcpyAsync(greedy_out, r.ids + (size_t)K * B * Tc, (size_t)B * Tc * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int check_ready(s2vt_handle* h) {
    if (!h) return S2VT_EINVAL;
    if (!h->bound) return h->fail(S2VT_ESTATE, "s2vt_bind has not been called");
    if (!h->fresh) return h->fail(S2VT_ESTATE, "parameters changed: call s2vt_refresh first");
    return 0;
}

extern "C" int s2vt_rollout(s2vt_handle* h, const float* video, int B, int K, uint64_t seed, uint32_t row_base, int32_t* sampled_out, int32_t* greedy_out,
                            s2vt_stream st) {
    TRY(check_ready(h));
    if (B <= 0 || K < 0 || !video) return h->fail(S2VT_EINVAL, "bad rollout arguments");
    return h->cfg.precision == S2VT_PREC_BF16 ? rollout_impl<bf16>(h, (cudaStream_t)st, video, B, K, seed, row_base, sampled_out, greedy_out)
                                              : rollout_impl<float>(h, (cudaStream_t)st, video, B, K, seed, row_base, sampled_out, greedy_out);
}
extern "C" int s2vt_greedy(s2vt_handle* h, const float* video, int B, int32_t* ids_out, s2vt_stream st) {
    if (!ids_out) return S2VT_EINVAL;
    return s2vt_rollout(h, video, B, 0, 0, 0, nullptr, ids_out, st);
}
extern "C" int s2vt_caption_masks(s2vt_handle* h, const int32_t* ids, int N, float* mask_out, int32_t* lengths_out, s2vt_stream st) {
    if (!h || N <= 0) return S2VT_EINVAL;
    caption_mask_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)st>>>(ids, N, h->Tc, mask_out, lengths_out);
    KCHECK(h);
    return 0;
}

// =================================================================================================================
// teacher-forced forward / backward
// =================================================================================================================
template <typename T>
struct Train {
    typedef typename Fwd<T>::type F;     // forward values F, gradients T
    Front<T> f;
    F* out1d; float* G2x; F* h2_all; float* c2_all; F* gates2; F* out2d; float* logits;
    int *prev_tok, *target; float *ca, *cb, *cc, *logp, *sumlsm;
    // backward
    T* dlogits; float* dout2; T* dG2; float* dc2; float* dout1; float* dEmb; float* dh1; T* dG1; float* dc1; float* dimgF; T* dimgT_src;
    T *tA, *tB;   // transposed operand scratch (largest: [Vp, Mp])
    T *tA2, *tB2; // same for the LSTM1 chain on the side stream
    void *chain_f, *chain_b2, *chain_b1;   // per-step parameters of the persistent chains
    float* ws2_scratch; unsigned* ws2_flags;   // partial tiles / counters of the weights-stationary BPTT chain (> 128 rows)
    F* emb;
};

template <typename T>
static void plan_train(const s2vt_handle* h, Arena& a, int B, int N, bool backward, bool want_logits_only, Train<T>& p) {
    typedef typename Fwd<T>::type F;
    const int T_ = h->T, Tv = h->Tv, Tc = h->Tc, Hp = h->Hp, Gp = h->Gp, Vp = h->Vp, Ep = h->Ep, Dp = h->Dp;
    plan_front<T>(h, a, B, backward, p.f);
    p.out1d = a.take<F>((size_t)T_ * N * Hp);
    p.G2x = a.take<float>((size_t)T_ * N * Gp);
    p.h2_all = a.take<F>((size_t)(T_ + 1) * N * Hp);
    p.c2_all = a.take<float>((size_t)(T_ + 1) * N * Hp);
    p.gates2 = backward ? a.take<F>((size_t)T_ * N * Gp) : nullptr;
    p.out2d = a.take<F>((size_t)Tc * N * Hp);
    p.logits = a.take<float>((size_t)Tc * N * Vp);
    p.prev_tok = a.take<int>((size_t)Tc * N); p.target = a.take<int>((size_t)Tc * N);
    p.ca = a.take<float>((size_t)Tc * N); p.cb = a.take<float>((size_t)Tc * N); p.cc = a.take<float>((size_t)Tc * N);
    p.logp = a.take<float>((size_t)Tc * N); p.sumlsm = a.take<float>((size_t)Tc * N);
    p.chain_f = a.take<char>((size_t)(T_ + 1) * 512); p.chain_b2 = a.take<char>((size_t)(T_ + 1) * 512); p.chain_b1 = a.take<char>((size_t)(T_ + 1) * 512);
    if (!backward) return;
    const size_t MpD = ru(Tc * N, S2VT_PAD), Mp2 = ru(T_ * N, S2VT_PAD), Mp1 = ru(T_ * B, S2VT_PAD);
    p.dlogits = a.take<T>((size_t)Tc * N * Vp);
    p.dout2 = a.take<float>((size_t)Tc * N * Hp);
    p.dG2 = a.take<T>((size_t)T_ * N * Gp);
    p.dc2 = a.take<float>((size_t)N * Hp);
    p.dout1 = a.take<float>((size_t)T_ * N * Hp);
    p.dEmb = a.take<float>((size_t)Tc * N * Ep);
    p.dh1 = a.take<float>((size_t)T_ * B * Hp);
    p.dG1 = a.take<T>((size_t)T_ * B * Gp);
    p.dc1 = a.take<float>((size_t)B * Hp);
    p.dimgF = a.take<float>((size_t)Tv * B * Ep);
    p.dimgT_src = a.take<T>((size_t)Tv * B * Ep);
    p.emb = a.take<F>((size_t)Tc * N * Ep);
    p.ws2_scratch = N > 128 ? a.take<float>(tc::ws2_bwd_scratch_floats(Hp)) : nullptr;
    p.ws2_flags = a.take<unsigned>(256);
    size_t ta = (size_t)Hp * Mp2;                       // activations^T : at most [max(Hp,Ep,Dp), Mp2]
    if ((size_t)Dp * Mp1 > ta) ta = (size_t)Dp * Mp1;
    if ((size_t)Ep * MpD > ta) ta = (size_t)Ep * MpD;
    size_t tb = (size_t)Vp * MpD;                       // gradients^T  : [Vp, MpD] or [Gp, Mp2]
    if ((size_t)Gp * Mp2 > tb) tb = (size_t)Gp * Mp2;
    p.tA = a.take<T>(ta + 256);
    p.tB = a.take<T>(tb + 256);
    size_t ta2 = (size_t)(Dp > Hp ? Dp : Hp) * Mp1, tb2 = (size_t)Gp * Mp1;
    p.tA2 = a.take<T>(ta2 + 256);
    p.tB2 = a.take<T>(tb2 + 256);
}

template <typename TI, typename TO>
static int transpose(s2vt_handle* h, cudaStream_t st, const TI* src, int lds, int R, int C, TO* dst, int ldd, int rows_dst_padded) {
    // zero the destination (pads along the contraction dimension must be zero), then transpose the valid part
    CUDA_TRY(h, cudaMemsetAsync(dst, 0, (size_t)rows_dst_padded * ldd * sizeof(TO), st));
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    transpose_kernel<TI, TO><<<grid, block, 0, st>>>(src, lds, R, C, dst, ldd);
    KCHECK(h);
    return 0;
}

// mode 0: REINFORCE (needs rewards/base), mode 1: XE, mode 2: forward only
template <typename T>
static int train_impl(s2vt_handle* h, cudaStream_t st, int mode, const float* video, int B, const int32_t* captions, const float* mask, const float* rewards,
                      const float* base_line, int N, float norm, float grad_scale, int accumulate, float ls, float decay, uint64_t drop_seed,
                      uint32_t row_base, float* loss_out, float* logp_out, float* logits_out, const float* xe_colsum = nullptr, int xe_nglobal = 0) {
    typedef typename Fwd<T>::type F;     // forward operands / activations (fp16 in the bf16 mode); gradients stay T
    const int Tv = h->Tv, Tc = h->Tc, T_ = h->T, Hp = h->Hp, Gp = h->Gp, Vp = h->Vp, Ep = h->Ep, Dp = h->Dp;
    const int H = h->H, E = h->E, V = h->V, D = h->D, G = 4 * h->H;
    const bool backward = mode != 2;
    const float keep = (drop_seed == 0) ? 1.f : h->cfg.dropout_keep;
    if (N % B != 0) return h->fail(S2VT_EINVAL, "N (%d) must be a multiple of B (%d): row n uses video n %% B", N, B);
    Arena a(h->ws, h->ws_bytes);
    Train<T> p;
    plan_train<T>(h, a, B, N, true, false, p);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small: need %zu bytes", a.used);

    // ---------------- forward ----------------
    // LSTM1 (and the frame projection) depend on the video only: reuse the pass the preceding rollout already made
    if (!(backward && h->reuse_front && h->front_valid && h->front_B == B && h->front_video == video)) {
        h->front_valid = false;
        TRY(run_front<T>(h, st, video, B, p.f));
    }
    TRY(wait_late_weights(h, st));
    expand_dropout_kernel<F><<<T_ * N, 256, 0, st>>>(p.f.h1_all + (size_t)B * Hp, B, N, Hp, H, p.out1d, drop_seed, S2VT_STREAM_DROP1, row_base, keep);
    KCHECK(h);
    {
        typename EpiStore<F>::Params ep = {p.G2x, nullptr, Gp, nullptr, T_ * N, 0};
        TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, p.out1d, Hp, h->W2xT, Hp, T_ * N, Gp, Hp, ep)));
    }
    caption_tables_kernel<<<(Tc * N + 255) / 256, 256, 0, st>>>(captions, N, Tc, p.prev_tok, p.target); KCHECK(h);
    CUDA_TRY(h, cudaMemsetAsync(p.h2_all, 0, (size_t)N * Hp * sizeof(F), st));
    CUDA_TRY(h, cudaMemsetAsync(p.c2_all, 0, (size_t)N * Hp * sizeof(float), st));
    {
        StepChain<F, EpiLstmFwd<F>> ch;
        ch.A = p.h2_all; ch.lda = Hp; ch.a_total_rows = (T_ + 1) * N; ch.a_row0 = 0; ch.a_row_stride = N;
        ch.B = (const F*)h->W2hT; ch.ldb = Hp; ch.M = N; ch.N = Gp; ch.K = Hp;
        for (int t = 0; t < T_; ++t) {
            typename EpiLstmFwd<F>::Params ep;
            memset(&ep, 0, sizeof ep);
            ep.M = N; ep.Hp = Hp; ep.bias = h->b2_p; ep.add0 = p.G2x + (size_t)t * N * Gp;
            if (t >= Tv) { ep.add1 = reinterpret_cast<const F*>(h->Etab); ep.tok = p.prev_tok + (size_t)(t - Tv) * N; ep.hdrop_out = p.out2d + (size_t)(t - Tv) * N * Hp; }
            ep.c_prev = p.c2_all + (size_t)t * N * Hp; ep.c_out = p.c2_all + (size_t)(t + 1) * N * Hp;
            ep.h_out = p.h2_all + (size_t)(t + 1) * N * Hp;
            ep.gates_out = p.gates2 ? p.gates2 + (size_t)t * N * Gp : nullptr;
            ep.seed = drop_seed; ep.stream = S2VT_STREAM_DROP2; ep.step = (uint32_t)t; ep.row_base = row_base; ep.keep = keep;
            ch.eps.push_back(ep);
        }
        TRY((run_chain<F, EpiLstmFwd<F>>(h, st, ch, p.chain_f)));
    }
    {   // logits for all decode steps at once (:163 / :286)
        typename EpiStore<F>::Params ep = {p.logits, nullptr, Vp, h->bo_p, Tc * N, 0};
        TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, p.out2d, Hp, h->WoT, Hp, Tc * N, Vp, Hp, ep)));
    }
    if (logits_out)
        CUDA_TRY(h, cudaMemcpy2DAsync(logits_out, (size_t)V * sizeof(float), p.logits, (size_t)Vp * sizeof(float), (size_t)V * sizeof(float), (size_t)Tc * N,
                                      cudaMemcpyDeviceToDevice, st));
    if (!backward) {
        softmax_rows_kernel<T><<<Tc * N, ROW_THREADS, 0, st>>>(p.logits, Vp, V, Vp, p.target, nullptr, nullptr, nullptr, p.logp, p.sumlsm, (T*)nullptr, logp_out,
                                                               N, Tc);
        KCHECK(h);
        return 0;
    }
    // ---------------- loss coefficients ----------------
    float* nrm = h->scal;   // scal[0] = norm used by the objective, scal[1] = sum(mask) of this call
    sum_kernel<<<1, 256, 0, st>>>(mask, N * Tc, h->scal + 1); KCHECK(h);
    set_norm_kernel<<<1, 1, 0, st>>>(nrm, h->scal + 1, norm); KCHECK(h);
    if (!accumulate) CUDA_TRY(h, cudaMemsetAsync(h->grads, 0, (h->P + 8) * sizeof(float), st));
    loss_coef_kernel<<<(Tc * N + 255) / 256, 256, 0, st>>>(mode, mask, rewards, base_line, N, Tc, nrm, grad_scale, ls, V, p.ca, p.cb, p.cc, xe_colsum, xe_nglobal); KCHECK(h);
    softmax_rows_kernel<T><<<Tc * N, ROW_THREADS, 0, st>>>(p.logits, Vp, V, Vp, p.target, p.ca, p.cb, p.cc, p.logp, p.sumlsm, p.dlogits, logp_out, N, Tc);
    KCHECK(h);
    loss_reduce_kernel<<<1, 256, 0, st>>>(mode, p.logp, p.sumlsm, mask, rewards, base_line, N, Tc, nrm, ls, V, h->scal + 4, nullptr, xe_colsum, xe_nglobal); KCHECK(h);
    if (loss_out) {   // loss_out[0] (+)= grad_scale * objective, so a sequence of accumulate calls yields the mixed loss
        if (!accumulate) CUDA_TRY(h, cudaMemsetAsync(loss_out, 0, sizeof(float), st));
        add_scaled_scalar_kernel<<<1, 1, 0, st>>>(loss_out, h->scal + 4, grad_scale); KCHECK(h);
    }
    add_scaled_scalar_kernel<<<1, 1, 0, st>>>(h->grads + h->P + 1, h->scal + 4, grad_scale); KCHECK(h);   // aux[1] = loss (summed by the DP allreduce)
    if (mode == 0) { add_scaled_scalar_kernel<<<1, 1, 0, st>>>(h->grads + h->P + 2, h->scal + 1, 1.f); KCHECK(h); }   // aux[2] = sum(mask)

    // ---------------- backward ----------------
    const int MD = Tc * N, M2 = T_ * N, M1 = T_ * B, ME = Tv * B;
    const int MpD = ru(MD, S2VT_PAD), Mp2 = ru(M2, S2VT_PAD), Mp1 = ru(M1, S2VT_PAD), MpE = ru(ME, S2VT_PAD);
    {   // dout2 = dlogits . embed_word_W^T
        typename EpiStore<T>::Params ep = {p.dout2, nullptr, Hp, nullptr, MD, 0};
        TRY((gemm<T, CfgBig, EpiStore<T>>(h, st, p.dlogits, Vp, h->Wo, Vp, MD, Hp, Vp, ep)));
    }
    // ---- side stream 1/2: the vocabulary-projection weight gradient only needs dlogits / out2, so it runs (large GEMM, fills
    //      the idle SMs) while the main stream walks the latency-bound LSTM2 BPTT chain.
    TRY(ensure_side(h));
    cudaStream_t s2 = (h->overlap & 2) ? h->side : st;
    cudaStream_t s3 = (h->overlap & 4) ? h->side : st;   // stream of the LSTM1 backward chain
    CUDA_TRY(h, cudaMemsetAsync(h->gbar, 0, sizeof(unsigned), st));   // the chain-progress watcher on the side stream must never see the count of an earlier chain
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(h, cudaStreamWaitEvent(s2, h->ev_fork, 0));
    // d embed_word_W = out2^T . dlogits ; d embed_word_b = column sums.  overlap bit 6: not beside the LSTM2 BPTT chain (whose steps it slows from 12 to
    // 15 us) but later on the caller's stream, among the LSTM2 weight gradients that run beside the LSTM1 chain.
    const bool dwo_late = (h->overlap & 64) != 0;
    auto dwo = [&](cudaStream_t sx) -> int {
        EpiGradStore::Params ep = {h->G_(h->iWo), V, H, V, 0, 1.f};
        TRY((wgrad<T, F>(h, sx, p.out2d, Hp, Hp, p.dlogits, Vp, Vp, MD, ep, p.tA, p.tB, H)));
        TRY(bias_grad<T>(h, sx, p.dlogits, Vp, Vp, MD, V, 0, h->G_(h->ibo)));
        return 0;
    };
    if (!dwo_late) TRY(dwo(s2));
    CUDA_TRY(h, cudaEventRecord(h->ev_join, s2));
    // Data-parallel hook: in the REINFORCE objective nothing touches these two gradients again (no weight decay, no accumulation
    // pass), so a caller may start their all-reduce now, under the BPTT chains (s2vt_grad_segment_ready).
    h->wo_grad_early = mode == 0 && !accumulate && grad_scale == 1.f && !dwo_late;
    if (h->wo_grad_early) CUDA_TRY(h, cudaEventRecord(h->ev_wo, s2));
    h->seg_ready = h->wo_grad_early ? 1u : 0u;
    // LSTM2 BPTT
    CUDA_TRY(h, cudaMemsetAsync(p.dc2, 0, (size_t)N * Hp * sizeof(float), st));
    int gate_ncta = 0;
    {
        StepChain<T, EpiLstmBwd<T, F>> ch;    // step s of the chain is time t = T-2-s; its A operand is dG2 of time t+1
        ch.A = p.dG2; ch.lda = Gp; ch.a_total_rows = T_ * N; ch.a_row0 = (T_ - 1) * N; ch.a_row_stride = -N;
        ch.B = (const T*)h->W2h; ch.ldb = Gp; ch.M = N; ch.N = Hp; ch.K = Gp;
        ch.ws2_scratch = p.ws2_scratch; ch.ws2_flags = p.ws2_flags; ch.ws2_flags_cap = 256;
        for (int t = T_ - 1; t >= 0; --t) {
            LstmBwdArgs b;
            memset(&b, 0, sizeof b);
            b.M = N; b.Hp = Hp;
            b.dh_ext = t >= Tv ? p.dout2 + (size_t)(t - Tv) * N * Hp : nullptr;
            b.gates = p.gates2 + (size_t)t * N * Gp; b.c_prev = p.c2_all + (size_t)t * N * Hp; b.c_new = p.c2_all + (size_t)(t + 1) * N * Hp;
            b.dc = p.dc2; b.seed = drop_seed; b.stream = S2VT_STREAM_DROP2; b.step = (uint32_t)t; b.row_base = row_base; b.keep = keep;
            T* dg = p.dG2 + (size_t)t * N * Gp;
            if (t == T_ - 1) {
                lstm_bwd_elem_kernel<T, F><<<(N * Hp + 255) / 256, 256, 0, st>>>(b, dg); KCHECK(h);
            } else {
                typename EpiLstmBwd<T, F>::Params ep = {b, dg};
                ch.eps.push_back(ep);
            }
        }
        TRY((run_chain<T, EpiLstmBwd<T, F>>(h, st, ch, p.chain_b2)));
        gate_ncta = ch.gate_ncta;
    }
    // ---- side stream 2/2, gated (overlap bit 7): the chain produces dG2 from the last time step down; once it has passed time t_gate, the rows of the times
    //      >= t_gate are final, and what consumes them -- their part of dout1 = dG2 . W2[x rows]^T and of the two large LSTM2 weight gradients -- runs on the SMs
    //      the 96-CTA chain leaves idle, behind a watcher of the chain's grid-barrier counter, instead of after the chain.  The rest (times < t_gate) follows on
    //      the caller's stream as before; the weight gradients are read-modify-write sums, so the two parts simply add up.
    int t_gate = T_;
    {
        static const float gate_frac = getenv("S2VT_GATE_FRAC") ? (float)atof(getenv("S2VT_GATE_FRAC")) : 0.4f;
        if ((h->overlap & 128) && s2 == h->side && s2 != st && gate_ncta > 0 && !h->prof && T_ >= 16 && gate_frac > 0.f) {
            t_gate = T_ - (int)(T_ * gate_frac);
            if (t_gate < 1) t_gate = 1;
        }
    }
    if (t_gate < T_) {
        const int rowsA = (T_ - t_gate) * N;
        const size_t off = (size_t)t_gate * N;
        tc::chain_progress_wait_kernel<<<1, 32, 0, s2>>>(h->gbar, (unsigned)(T_ - 1 - t_gate) * (unsigned)gate_ncta); KCHECK(h);   // chain step s handles time T-2-s
        typename EpiStore<T>::Params ep = {p.dout1 + off * Hp, nullptr, Hp, nullptr, rowsA, 0};
        TRY((gemm<T, CfgBig, EpiStore<T>>(h, s2, p.dG2 + off * Gp, Gp, h->W2x, Gp, rowsA, Hp, Gp, ep)));
        CUDA_TRY(h, cudaEventRecord(h->ev_gate, s2));
        float* gW2 = h->G_(h->iW2);
        EpiGradStore::Params e1 = {gW2, G, H, G, H, 1.f};
        TRY((wgrad<T, F>(h, s2, p.out1d + off * Hp, Hp, Hp, p.dG2 + off * Gp, Gp, Gp, rowsA, e1, p.tA, p.tB, H)));
        EpiGradStore::Params e3 = {gW2 + (size_t)(H + E) * G, G, H, G, H, 1.f};
        TRY((wgrad<T, F>(h, s2, p.h2_all + off * Hp, Hp, Hp, p.dG2 + off * Gp, Gp, Gp, rowsA, e3, p.tA, p.tB, H)));
        CUDA_TRY(h, cudaEventRecord(h->ev_join, s2));     // tA / tB are free after these
    }
    const int MB2 = t_gate * N;     // rows (times < t_gate) handled on the caller's stream
    {   // gradient flowing into LSTM1's (dropped, shared-per-video) output
        typename EpiStore<T>::Params ep = {p.dout1, nullptr, Hp, nullptr, MB2, 0};
        TRY((gemm<T, CfgBig, EpiStore<T>>(h, st, p.dG2, Gp, h->W2x, Gp, MB2, Hp, Gp, ep)));
        if (t_gate < T_) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_gate, 0));
        reduce_dropout_kernel<<<M1, 256, 0, st>>>(p.dout1, B, N, Hp, p.dh1, drop_seed, S2VT_STREAM_DROP1, row_base, keep); KCHECK(h);
    }
    // ---- fork: the LSTM1 chain (B rows, few CTAs per step) runs on the side stream while the main stream computes the
    //      LSTM2 / vocabulary weight gradients (large GEMMs); joined before returning.
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join, 0));   // dWo done: tA / tB are free again
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(h, cudaStreamWaitEvent(s3, h->ev_fork, 0));
    // LSTM1 BPTT over the B shared rows (side stream)
    CUDA_TRY(h, cudaMemsetAsync(p.dc1, 0, (size_t)B * Hp * sizeof(float), s3));
    {
        StepChain<T, EpiLstmBwd<T, F>> ch;
        ch.A = p.dG1; ch.lda = Gp; ch.a_total_rows = T_ * B; ch.a_row0 = (T_ - 1) * B; ch.a_row_stride = -B;
        ch.B = (const T*)h->W1h; ch.ldb = Gp; ch.M = B; ch.N = Hp; ch.K = Gp;
        for (int t = T_ - 1; t >= 0; --t) {
            LstmBwdArgs b;
            memset(&b, 0, sizeof b);
            b.M = B; b.Hp = Hp; b.dh_ext = p.dh1 + (size_t)t * B * Hp;
            b.gates = p.f.gates1 + (size_t)t * B * Gp; b.c_prev = p.f.c1_all + (size_t)t * B * Hp; b.c_new = p.f.c1_all + (size_t)(t + 1) * B * Hp;
            b.dc = p.dc1; b.keep = 1.f;
            T* dg = p.dG1 + (size_t)t * B * Gp;
            if (t == T_ - 1) {
                lstm_bwd_elem_kernel<T, F><<<(B * Hp + 255) / 256, 256, 0, s3>>>(b, dg); KCHECK(h);
            } else {
                typename EpiLstmBwd<T, F>::Params ep = {b, dg};
                ch.eps.push_back(ep);
            }
        }
        TRY((run_chain<T, EpiLstmBwd<T, F>>(h, s3, ch, p.chain_b1)));
    }
    {   // LSTM1 kernel / bias gradients (side stream, own transpose scratch)
        float* gW1 = h->G_(h->iW1);
        TRY(bias_grad<T>(h, s3, p.dG1, Gp, Gp, M1, 0, H, h->G_(h->ib1)));
        EpiGradStore::Params e2 = {gW1 + (size_t)E * G, G, H, G, H, 1.f};
        TRY((wgrad<T, F>(h, s3, p.f.h1_all, Hp, Hp, p.dG1, Gp, Gp, M1, e2, p.tA2, p.tB2, H)));     // h1 before step t = h1_all[t]
        // frame-embedding rows: encoder steps only
        EpiGradStore::Params e1 = {gW1, G, E, G, H, 1.f};
        TRY((wgrad<T, F>(h, s3, p.f.img, Ep, Ep, p.dG1, Gp, Gp, ME, e1, p.tA2, p.tB2, E)));
    }
    {   // frame projection gradients: dimg = dG1[enc] . W1[x rows]^T ; dWe = X^T . dimg ; dbe = column sums (side stream)
        typename EpiStore<T>::Params ep = {p.dimgF, p.dimgT_src, Ep, nullptr, ME, 0};
        TRY((gemm<T, CfgBig, EpiStore<T>>(h, s3, p.dG1, Gp, h->W1x, Gp, ME, Ep, Gp, ep)));
        TRY(bias_grad<T>(h, s3, p.dimgT_src, Ep, Ep, ME, E, 0, h->G_(h->ibe)));
        EpiGradStore::Params e = {h->G_(h->iWe), E, D, E, 0, 1.f};
        TRY((wgrad<T, F>(h, s3, p.f.Xc, Dp, Dp, p.dimgT_src, Ep, Ep, ME, e, p.tA2, p.tB2, D)));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_join, s3));
    // ---- main stream meanwhile: embedding and LSTM2 weight gradients (and, with overlap bit 6, the vocabulary-projection ones)
    if (dwo_late) TRY(dwo(st));
    {   // word-embedding gradient
        typename EpiStore<T>::Params ee = {p.dEmb, nullptr, Ep, nullptr, MD, 0};
        TRY((gemm<T, CfgBig, EpiStore<T>>(h, st, p.dG2 + (size_t)Tv * N * Gp, Gp, h->W2e, Gp, MD, Ep, Gp, ee)));
        scatter_emb_grad_kernel<<<MD, 128, 0, st>>>(p.dEmb, Ep, p.prev_tok, MD, E, h->G_(h->iWemb), h->grads + h->P); KCHECK(h);   // aux[0] = slice square norm (R6)
        if (h->wo_grad_early) { CUDA_TRY(h, cudaEventRecord(h->ev_seg[0], st)); h->seg_ready |= 2u; }      // d Wemb is final (the aux slot travels with the remainder)
    }
    {   // LSTM2 kernel / bias gradients: [out1 ; emb ; h2]^T . dG2
        float* gW2 = h->G_(h->iW2);
        TRY(bias_grad<T>(h, st, p.dG2, Gp, Gp, M2, 0, H, h->G_(h->ib2)));
        EpiGradStore::Params e1 = {gW2, G, H, G, H, 1.f};
        TRY((wgrad<T, F>(h, st, p.out1d, Hp, Hp, p.dG2, Gp, Gp, MB2, e1, p.tA, p.tB, H)));            // times < t_gate (all of them without the gated part)
        EpiGradStore::Params e3 = {gW2 + (size_t)(H + E) * G, G, H, G, H, 1.f};
        TRY((wgrad<T, F>(h, st, p.h2_all, Hp, Hp, p.dG2, Gp, Gp, MB2, e3, p.tA, p.tB, H)));           // h2 before step t = h2_all[t]
        // embedding rows: decode steps only
        gather_rows_kernel<F><<<MD, 128, 0, st>>>((const F*)h->WembC, Ep, p.prev_tok, MD, p.emb); KCHECK(h);
        EpiGradStore::Params e2 = {gW2 + (size_t)H * G, G, E, G, H, 1.f};
        TRY((wgrad<T, F>(h, st, p.emb, Ep, Ep, p.dG2 + (size_t)Tv * N * Gp, Gp, Gp, MD, e2, p.tA, p.tB, E)));
        if (h->wo_grad_early) { CUDA_TRY(h, cudaEventRecord(h->ev_seg[1], st)); h->seg_ready |= 4u; }      // d LSTM2 weights / biases are final
    }
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join, 0));   // join
    if (mode == 1 && decay > 0.f) {   // Q4: L2 on every variable without 'bias' in its name (the LSTM '/biases' only)
        CUDA_TRY(h, cudaMemsetAsync(h->sq + 3, 0, sizeof(double), st));
        for (size_t i = 0; i < h->vars.size(); ++i) {
            if ((int)i == h->ib1 || (int)i == h->ib2) continue;
            if ((int)i == h->iAW || (int)i == h->iAb) continue;   // attribute head is not part of build_model's graph here
            const Var& v = h->vars[i];
            add_decay_kernel<<<148 * 4, 256, 0, st>>>(h->grads + v.off, h->params + v.off, v.count(), decay * grad_scale); KCHECK(h);
            sumsq_kernel<<<148 * 2, 256, 0, st>>>(h->params + v.off, v.count(), h->sq + 3); KCHECK(h);
        }
        if (loss_out) { xe_total_kernel<<<1, 1, 0, st>>>(loss_out, h->sq, decay, grad_scale); KCHECK(h); }
    } else if (mode == 1 && loss_out) {
        CUDA_TRY(h, cudaMemsetAsync(loss_out + 1, 0, sizeof(float), st));   // no weight-decay part (decay 0: e.g. every rank but one of a sharded batch)
    }
    return 0;
}

#define DISPATCH(h, call_bf16, call_f32) ((h)->cfg.precision == S2VT_PREC_BF16 ? (call_bf16) : (call_f32))

extern "C" int s2vt_teacher_forward(s2vt_handle* h, const float* video, int B, const int32_t* captions, int N, uint64_t drop_seed, uint32_t row_base,
                                    float* logp_out, float* logits_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!video || !captions || B <= 0 || N <= 0) return h->fail(S2VT_EINVAL, "bad teacher_forward arguments");
    cudaStream_t s = (cudaStream_t)st;
    return DISPATCH(h, (train_impl<bf16>(h, s, 2, video, B, captions, nullptr, nullptr, nullptr, N, 0.f, 1.f, 0, 0.f, 0.f, drop_seed, row_base, nullptr, logp_out, logits_out)),
                    (train_impl<float>(h, s, 2, video, B, captions, nullptr, nullptr, nullptr, N, 0.f, 1.f, 0, 0.f, 0.f, drop_seed, row_base, nullptr, logp_out, logits_out)));
}

extern "C" int s2vt_rl_backward(s2vt_handle* h, const float* video, int B, const int32_t* captions, const float* mask, const float* rewards,
                                const float* base_line, int N, float norm, float grad_scale, int accumulate, uint64_t drop_seed, uint32_t row_base,
                                float* loss_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!video || !captions || !mask || !rewards || !base_line || B <= 0 || N <= 0) return h->fail(S2VT_EINVAL, "bad rl_backward arguments");
    cudaStream_t s = (cudaStream_t)st;
    return DISPATCH(h, (train_impl<bf16>(h, s, 0, video, B, captions, mask, rewards, base_line, N, norm, grad_scale, accumulate, 0.f, 0.f, drop_seed, row_base, loss_out, nullptr, nullptr)),
                    (train_impl<float>(h, s, 0, video, B, captions, mask, rewards, base_line, N, norm, grad_scale, accumulate, 0.f, 0.f, drop_seed, row_base, loss_out, nullptr, nullptr)));
}

extern "C" int s2vt_xe_backward(s2vt_handle* h, const float* video, int B, const int32_t* captions, const float* mask, int N, float label_smoothing,
                                float decay, float norm, float grad_scale, int accumulate, uint64_t drop_seed, uint32_t row_base, float* loss_out,
                                s2vt_stream st) {
    TRY(check_ready(h));
    if (!video || !captions || !mask || B <= 0 || N <= 0) return h->fail(S2VT_EINVAL, "bad xe_backward arguments");
    cudaStream_t s = (cudaStream_t)st;
    return DISPATCH(h, (train_impl<bf16>(h, s, 1, video, B, captions, mask, nullptr, nullptr, N, norm, grad_scale, accumulate, label_smoothing, decay, drop_seed, row_base, loss_out, nullptr, nullptr)),
                    (train_impl<float>(h, s, 1, video, B, captions, mask, nullptr, nullptr, N, norm, grad_scale, accumulate, label_smoothing, decay, drop_seed, row_base, loss_out, nullptr, nullptr)));
}

// Data-parallel form of the XE objective: Q3 couples the rows of a batch (mean_b(CE_b) * sum_b mask[b, i]), so a rank that holds a
// shard needs the GLOBAL per-step mask sums and the global row count; with them the per-rank gradients and losses simply add up.
extern "C" int s2vt_xe_backward_sharded(s2vt_handle* h, const float* video, int B, const int32_t* captions, const float* mask, int N, float label_smoothing,
                                        float decay, float norm, const float* mask_colsum_global, int n_rows_global, float grad_scale, int accumulate,
                                        uint64_t drop_seed, uint32_t row_base, float* loss_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!video || !captions || !mask || !mask_colsum_global || B <= 0 || N <= 0 || n_rows_global < N || !(norm > 0.f))
        return h->fail(S2VT_EINVAL, "bad xe_backward_sharded arguments (the global sum(mask), its per-step sums and the global row count are required)");
    cudaStream_t s = (cudaStream_t)st;
    return DISPATCH(h, (train_impl<bf16>(h, s, 1, video, B, captions, mask, nullptr, nullptr, N, norm, grad_scale, accumulate, label_smoothing, decay, drop_seed, row_base, loss_out, nullptr, nullptr, mask_colsum_global, n_rows_global)),
                    (train_impl<float>(h, s, 1, video, B, captions, mask, nullptr, nullptr, N, norm, grad_scale, accumulate, label_smoothing, decay, drop_seed, row_base, loss_out, nullptr, nullptr, mask_colsum_global, n_rows_global)));
}

// ---- attribute head (config 4) --------------------------------------------------------------------------------------
template <typename T>
static int attribute_impl(s2vt_handle* h, cudaStream_t st, const float* video, int B, const float* labels, float grad_scale, float* loss_out) {
    typedef typename Fwd<T>::type F;
    const int A = h->A, Ap = h->Ap, D = h->D, Dp = h->Dp, Bp = ru(B, S2VT_PAD);
    h->front_valid = false;
    TRY(wait_late_weights(h, st));
    Arena a(h->ws, h->ws_bytes);
    F* pooled = a.take<F>((size_t)B * Dp);
    float* z = a.take<float>((size_t)B * Ap);
    T* dz = a.take<T>((size_t)B * Ap);
    T* pooledT = a.take<T>((size_t)Dp * Bp);
    T* dzT = a.take<T>((size_t)Ap * Bp);
    float* battr = a.take<float>(Ap);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small: need %zu bytes", a.used);
    mean_frames_kernel<F><<<B, 256, 0, st>>>(video, h->Tv, D, Dp, pooled); KCHECK(h);
    CUDA_TRY(h, cudaMemsetAsync(battr, 0, Ap * sizeof(float), st));
    CUDA_TRY(h, cudaMemcpyAsync(battr, h->P_(h->iAb), A * sizeof(float), cudaMemcpyDeviceToDevice, st));
    typename EpiStore<F>::Params ep = {z, nullptr, Ap, battr, B, 0};
    TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, pooled, Dp, h->attrWT, Dp, B, Ap, Dp, ep)));
    sigmoid_ce_kernel<T><<<1, 256, 0, st>>>(z, Ap, labels, B, A, Ap, grad_scale, dz, h->scal + 8); KCHECK(h);
    if (loss_out) CUDA_TRY(h, cudaMemcpyAsync(loss_out, h->scal + 8, sizeof(float), cudaMemcpyDeviceToDevice, st));
    EpiGradStore::Params eg = {h->G_(h->iAW), A, D, A, 0, 1.f};
    TRY((wgrad<T, F>(h, st, pooled, Dp, Dp, dz, Ap, Ap, B, eg, pooledT, dzT, D)));
    TRY(bias_grad<T>(h, st, dz, Ap, Ap, B, A, 0, h->G_(h->iAb)));
    return 0;
}

extern "C" int s2vt_attribute_backward(s2vt_handle* h, const float* video, int B, const float* labels, float grad_scale, float* loss_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!h->A) return h->fail(S2VT_ESTATE, "handle was created without an attribute head");
    if (!video || !labels || B <= 0) return h->fail(S2VT_EINVAL, "bad attribute_backward arguments");
    return DISPATCH(h, attribute_impl<bf16>(h, (cudaStream_t)st, video, B, labels, grad_scale, loss_out),
                    attribute_impl<float>(h, (cudaStream_t)st, video, B, labels, grad_scale, loss_out));
}

// ---- optimiser ----------------------------------------------------------------------------------------------------
extern "C" int s2vt_grad_segment_ready(s2vt_handle* h, int segment, s2vt_stream stream, int64_t* offset, int64_t* count) {
    if (!h || !h->bound) return S2VT_ESTATE;
    if (segment < 0 || segment > 2 || !offset || !count) return h->fail(S2VT_EINVAL, "segments: 0 (embed_word_W, embed_word_b), 1 (Wemb), 2 (LSTM2 weights, biases)");
    if (!(h->seg_ready & (1u << segment))) return h->fail(S2VT_ESTATE, "the last backward call did not finish this segment early");
    cudaEvent_t ev = segment == 0 ? h->ev_wo : h->ev_seg[segment - 1];
    CUDA_TRY(h, cudaStreamWaitEvent((cudaStream_t)stream, ev, 0));
    const int first = segment == 0 ? h->iWo : (segment == 1 ? h->iWemb : h->iW2), last = segment == 0 ? h->ibo : (segment == 1 ? h->iWemb : h->ib2);
    *offset = (int64_t)h->vars[first].off;
    *count = (int64_t)(h->vars[last].off + h->vars[last].count() - h->vars[first].off);
    h->seg_ready &= ~(1u << segment);      // one hand-out per backward call
    if (segment == 0) h->wo_grad_early = false;
    return S2VT_OK;
}

extern "C" int s2vt_optimizer_step(s2vt_handle* h, float lr, float clip_norm, int64_t step, int flags, float* gnorm_out, s2vt_stream st_) {
    const int wemb_slice_norm = flags & 1, normalize = (flags >> 1) & 1;
    if (!h || !h->bound) return S2VT_ESTATE;
    if (step < 1) return h->fail(S2VT_EINVAL, "Adam step is 1-based");
    if (h->opt_sharded) return h->fail(S2VT_ESTATE, "the Adam slots are sharded over the ranks (s2vt_peer_optimizer_step): call s2vt_peer_gather_state first");
    cudaStream_t st = (cudaStream_t)st_;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    TRY(wait_late_weights(h, st));   // a previous refresh may still be reading the parameters on the side stream
    CUDA_TRY(h, cudaMemsetAsync(h->sq, 0, 3 * sizeof(double), st));
    sumsq_kernel<<<148 * 4, 256, 0, st>>>(h->grads, h->P, h->sq); KCHECK(h);
    const Var& we = h->vars[h->iWemb];
    sumsq_kernel<<<148, 256, 0, st>>>(h->grads + we.off, we.count(), h->sq + 1); KCHECK(h);
    copy_slice_norm_kernel<<<1, 1, 0, st>>>(h->grads + h->P, h->sq + 2); KCHECK(h);
    double lr_t = (double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step));
    adam_kernel<<<148 * 8, 256, 0, st>>>(h->params, h->grads, h->adam_m, h->adam_v, h->P, h->sq, wemb_slice_norm, normalize, clip_norm, (float)lr_t, b1, b2, eps, gnorm_out);
    KCHECK(h);
    h->fresh = false;
    return s2vt_refresh(h, st_);
}

// ---- data-parallel gradient exchange over NVLink peer memory (peer.cuh) -------------------------------------------------------
static int peer_close(s2vt_handle* h) {
    for (void*& p : h->peer_opened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    h->peer_rank = -1; h->peer_world = 0;
    return 0;
}
extern "C" int s2vt_peer_export(s2vt_handle* h, unsigned char* state_handle, int64_t* state_offset, unsigned char* comm_handle) {
    if (!h || !h->bound) return S2VT_ESTATE;
    if (!state_handle || !state_offset || !comm_handle) return S2VT_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == S2VT_PEER_HANDLE_BYTES, "handle size");
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range = nullptr;
    if (!range) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return h->fail(S2VT_ECUDA, "cuMemGetAddressRange is not available");
        range = (RangeFn)p;
    }
    CUdeviceptr base = 0; size_t size = 0;
    if (range(&base, &size, (CUdeviceptr)h->state) != CUDA_SUCCESS) return h->fail(S2VT_ECUDA, "the state block is not a device allocation");
    cudaIpcMemHandle_t hs, hc;
    CUDA_TRY(h, cudaIpcGetMemHandle(&hs, (void*)base));    // fails for memory that cannot be shared (e.g. expandable segments): the caller falls back to NCCL
    if (!h->peer_comm) {
        CUDA_TRY(h, cudaMalloc(&h->peer_comm, (size_t)2 << 20));     // its own 2 MB block: never part of an allocation that is exported twice
        CUDA_TRY(h, cudaMemset(h->peer_comm, 0, (size_t)2 << 20));
    }
    CUDA_TRY(h, cudaIpcGetMemHandle(&hc, h->peer_comm));
    memcpy(state_handle, &hs, sizeof hs); memcpy(comm_handle, &hc, sizeof hc);
    *state_offset = (int64_t)((CUdeviceptr)h->state - base);
    return S2VT_OK;
}
extern "C" int s2vt_peer_connect(s2vt_handle* h, int rank, int world, const unsigned char* state_handles, const int64_t* state_offsets, const unsigned char* comm_handles) {
    if (!h || !h->bound || !h->peer_comm) return S2VT_ESTATE;
    if (world < 2 || world > peer::MAXW || rank < 0 || rank >= world || !state_handles || !state_offsets || !comm_handles) return S2VT_EINVAL;
    peer_close(h);
    for (int q = 0; q < world; ++q) {
        if (q == rank) { h->peer_s[q] = h->state; h->peer_c[q] = h->peer_comm; continue; }
        cudaIpcMemHandle_t hs, hc;
        memcpy(&hs, state_handles + (size_t)q * sizeof hs, sizeof hs); memcpy(&hc, comm_handles + (size_t)q * sizeof hc, sizeof hc);
        void *ps = nullptr, *pc = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ps, hs, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) { h->peer_opened[2 * q] = ps; e = cudaIpcOpenMemHandle(&pc, hc, cudaIpcMemLazyEnablePeerAccess); }
        if (e != cudaSuccess) { (void)cudaGetLastError(); peer_close(h); return h->fail(S2VT_ECUDA, "peer exchange: cannot map rank %d's memory: %s", q, cudaGetErrorString(e)); }
        h->peer_opened[2 * q + 1] = pc;
        h->peer_s[q] = (char*)ps + state_offsets[q];
        h->peer_c[q] = pc;
    }
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    CUDA_TRY(h, cudaFuncSetAttribute(peer::peer_allreduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, peer::TMA_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(peer::peer_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, peer::TMA_SMEM));
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer::peer_step_kernel, 512, peer::TMA_SMEM));
    if (per_sm < 1) return h->fail(S2VT_ECUDA, "peer exchange: kernel does not fit");
    h->peer_grid = sms;      // one CTA per SM, co-resident by construction (cooperative launch)
    h->peer_rank = rank; h->peer_world = world;
    return S2VT_OK;
}
extern "C" int s2vt_peer_disconnect(s2vt_handle* h) {
    if (!h) return S2VT_EINVAL;
    return peer_close(h);
}
static void peer_args(s2vt_handle* h, peer::Args& a, const float* block) {      // block: which per-rank block a.p points at (params / adam_m / adam_v)
    memset(&a, 0, sizeof a);
    const size_t goff = (size_t)((char*)h->grads - h->state), boff = (size_t)((const char*)block - h->state);
    for (int q = 0; q < h->peer_world; ++q) {
        a.g[q] = reinterpret_cast<float*>(h->peer_s[q] + goff);
        a.p[q] = reinterpret_cast<float*>(h->peer_s[q] + boff);
        a.comm[q] = reinterpret_cast<peer::Comm*>(h->peer_c[q]);
    }
    a.rank = h->peer_rank; a.world = h->peer_world; a.epoch = ++h->peer_epoch;
    a.P = h->P; a.n = h->P + 8; a.n4 = a.n / 4;
    a.slice4 = (a.n4 + a.world - 1) / a.world;
    static const int use_tma = getenv("S2VT_PEER_TMA") ? atoi(getenv("S2VT_PEER_TMA")) : 1;
    a.tma = use_tma && a.world <= peer::TMA_WMAX;
}
// grads (+ aux slots) <- sum over the ranks, identical bits on every rank; every rank must make the call (same order of calls everywhere)
extern "C" int s2vt_peer_allreduce(s2vt_handle* h, s2vt_stream st_) {
    if (!h || !h->bound || h->peer_world < 2) return S2VT_ESTATE;
    peer::Args a;
    peer_args(h, a, h->params);
    void* args[] = {&a};
    CUDA_TRY(h, cudaLaunchCooperativeKernel((const void*)peer::peer_allreduce_kernel, dim3(h->peer_grid), dim3(512), args, peer::TMA_SMEM, (cudaStream_t)st_));
    h->launches++;
    return S2VT_OK;
}
// s2vt_peer_allreduce + s2vt_optimizer_step in one kernel per rank: every rank reduces, clips and applies TF Adam to its own slice of the flat vector
// and the updated PARAMETERS are gathered instead of the gradients.  Arguments as s2vt_optimizer_step.
extern "C" int s2vt_peer_optimizer_step(s2vt_handle* h, float lr, float clip_norm, int64_t step, int flags, float* gnorm_out, s2vt_stream st_) {
    if (!h || !h->bound || h->peer_world < 2) return S2VT_ESTATE;
    if (step < 1) return h->fail(S2VT_EINVAL, "Adam step is 1-based");
    cudaStream_t st = (cudaStream_t)st_;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    TRY(wait_late_weights(h, st));   // a previous refresh may still be reading the parameters on the side stream
    peer::Args a;
    peer_args(h, a, h->params);
    const Var& we = h->vars[h->iWemb];
    a.m = h->adam_m; a.v = h->adam_v;
    a.wemb_lo = we.off; a.wemb_hi = we.off + we.count();
    a.use_slice_norm = flags & 1; a.normalize = (flags >> 1) & 1;
    a.clip = clip_norm; a.b1 = b1; a.b2 = b2; a.eps = eps; a.gnorm_out = gnorm_out;
    a.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step)));
    void* args[] = {&a};
    CUDA_TRY(h, cudaLaunchCooperativeKernel((const void*)peer::peer_step_kernel, dim3(h->peer_grid), dim3(512), args, peer::TMA_SMEM, st));
    h->launches++;
    h->opt_sharded = true;
    h->fresh = false;
    return s2vt_refresh(h, st_);
}
// Adam slots of all slices into this rank's blocks (collective).  After it the slots are whole again on every rank: checkpoints, s2vt_optimizer_step.
extern "C" int s2vt_peer_gather_state(s2vt_handle* h, s2vt_stream st_) {
    if (!h || !h->bound || h->peer_world < 2) return S2VT_ESTATE;
    for (int which = 0; which < 2; ++which) {
        peer::Args a;
        peer_args(h, a, which ? h->adam_v : h->adam_m);
        int phase0 = 0;
        void* args[] = {&a, &phase0};
        CUDA_TRY(h, cudaLaunchCooperativeKernel((const void*)peer::peer_gather_kernel, dim3(h->peer_grid), dim3(512), args, 0, (cudaStream_t)st_));
        h->launches++;
    }
    h->opt_sharded = false;
    return S2VT_OK;
}

// ---- workspace sizing ---------------------------------------------------------------------------------------------
template <typename T>
static size_t ws_bytes_impl(const s2vt_handle* h, int nv, int nr, int beam) {
    size_t best = 0;
    {   // rollout: K*B + B rows
        Arena a(nullptr, 0); Roll<T> r; plan_roll<T>(h, a, nv, nr + nv, r); if (a.used > best) best = a.used;
    }
    {   // training with de-duplicated videos, and with one video per row (the literal reference feed)
        Arena a(nullptr, 0); Train<T> p; plan_train<T>(h, a, nv, nr, true, false, p); if (a.used > best) best = a.used;
    }
    if (beam > 0) {
        Arena a(nullptr, 0); Roll<T> r; plan_roll<T>(h, a, nv, nv * beam, r);
        a.take<float>((size_t)8 * nv * beam * 64);   // beam bookkeeping, see beam.cuh
        if (a.used > best) best = a.used;
    }
    return best + 4096;
}
extern "C" size_t s2vt_workspace_bytes(const s2vt_handle* h, int n_videos, int n_rows, int beam) {
    if (!h || n_videos <= 0) return 0;
    if (n_rows < n_videos) n_rows = n_videos;
    return h->cfg.precision == S2VT_PREC_BF16 ? ws_bytes_impl<bf16>(h, n_videos, n_rows, beam) : ws_bytes_impl<float>(h, n_videos, n_rows, beam);
}

#include "beam.cuh"

// ---- isolated GEMM timing (scripts/gemm_shapes.py): one product of a given shape on scratch operands, through the same dispatch
// (and gemm_backend) as the engine's own products.  mn_major: C[M,N] = X^T . Y over K rows (the weight-gradient form).
extern "C" int s2vt_debug_gemm(s2vt_handle* h, int M, int N, int K, int mn_major, int fp32_out, s2vt_stream st_) {
    if (!h || !h->bound) return S2VT_ESTATE;
    if (h->cfg.precision != S2VT_PREC_BF16) return h->fail(S2VT_EINVAL, "debug_gemm: bf16 mode only");
    if (M <= 0 || N % 128 != 0 || (!mn_major && K % 128 != 0) || (mn_major && M % 128 != 0)) return h->fail(S2VT_EINVAL, "debug_gemm: N (and K, or M for mn_major) must be multiples of 128");
    cudaStream_t st = (cudaStream_t)st_;
    h->front_valid = false;      // the scratch operands overwrite the workspace (incl. a cached LSTM1 pass)
    Arena a(h->ws, h->ws_bytes);
    bf16* A = a.take<bf16>((size_t)(mn_major ? K : M) * (mn_major ? M : K));
    bf16* B = a.take<bf16>((size_t)(mn_major ? K : N) * (mn_major ? N : K));
    float* outF = a.take<float>((size_t)M * N);
    f16* outT = a.take<f16>((size_t)M * N);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small: need %zu bytes", a.used);
    static size_t zeroed_a = 0, zeroed_b = 0;
    if (zeroed_a != (size_t)M * K || zeroed_b != (size_t)N * K) {      // operands: zeros (timing only), once per shape
        CUDA_TRY(h, cudaMemsetAsync(A, 0, (size_t)M * K * 2, st));
        CUDA_TRY(h, cudaMemsetAsync(B, 0, (size_t)N * K * 2, st));
        zeroed_a = (size_t)M * K; zeroed_b = (size_t)N * K;
    }
    if (mn_major) {
        EpiGradStore::Params ep = {outF, N, M, N, 0, 1.f};
        bf16* none = nullptr;
        return wgrad<bf16, bf16>(h, st, A, M, M, B, N, N, K, ep, none, none, M);
    }
    if (fp32_out) {
        typename EpiStore<f16>::Params ep = {outF, nullptr, N, nullptr, M, 0};
        return gemm<f16, CfgBig, EpiStore<f16>>(h, st, A, K, B, K, M, N, K, ep);
    }
    typename EpiStore<f16>::Params ep = {nullptr, outT, N, nullptr, M, 0};
    return gemm<f16, CfgBig, EpiStore<f16>>(h, st, A, K, B, K, M, N, K, ep);
}

// ---- instrumentation ------------------------------------------------------------------------------------------------
extern "C" long long s2vt_launch_count(const s2vt_handle* h) { return h ? h->launches : 0; }
extern "C" int s2vt_profile(s2vt_handle* h, int enable) {
    if (!h) return S2VT_EINVAL;
    h->prof = enable != 0;
    return 0;
}
// Synchronises the device, aggregates the bracketed GEMM launches per class (0: batched GEMMs, 1: recurrent-step GEMMs)
// and clears the record list.
// Per-shape breakdown of the bracketed launches (does not clear): up to `cap` distinct (cls, M, N, K) rows.
extern "C" int s2vt_profile_shapes(s2vt_handle* h, int cap, int* cls, int* M, int* N, int* K, double* ms, long long* count, double* bytes,
                                   long long* launches) {
    if (!h) return S2VT_EINVAL;
    CUDA_TRY(h, cudaDeviceSynchronize());
    int n = 0;
    for (auto& r : h->prof_recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        int i = 0;
        for (; i < n; ++i) if (cls[i] == r.cls && M[i] == r.M && N[i] == r.N && K[i] == r.K) break;
        if (i == n) { if (n >= cap) continue; cls[n] = r.cls; M[n] = r.M; N[n] = r.N; K[n] = r.K; ms[n] = 0; count[n] = 0; if (bytes) bytes[n] = 0; if (launches) launches[n] = 0; ++n; }
        ms[i] += t; count[i] += r.count;
        if (bytes) bytes[i] += r.bytes;
        if (launches) launches[i] += r.launches;
    }
    return n;
}
extern "C" int s2vt_profile_read(s2vt_handle* h, double* ms_out, double* flops_out, long long* launches_out) {
    if (!h || !ms_out || !flops_out || !launches_out) return S2VT_EINVAL;
    CUDA_TRY(h, cudaDeviceSynchronize());
    for (int c = 0; c < 2; ++c) { ms_out[c] = 0; flops_out[c] = 0; flops_out[2 + c] = 0; launches_out[c] = 0; }
    for (auto& r : h->prof_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        ms_out[r.cls] += ms; flops_out[r.cls] += r.flops; flops_out[2 + r.cls] += r.bytes; launches_out[r.cls] += r.count;
        h->prof_pool.push_back(r.a); h->prof_pool.push_back(r.b);
    }
    h->prof_recs.clear();
    return 0;
}
