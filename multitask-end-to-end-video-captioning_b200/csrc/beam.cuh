// Batched on-device beam search (SURVEY K7 + K12).  Included by s2vt_api.cu (needs its forward plans).
//
// Restates the host loop of final_beam_search.py:248-294 (= e2e_beam_search.py:301-344) for B videos at once with
// per-video bookkeeping; semantics B1-B7 of SURVEY.md 3.4:
//   B1 only the best k of the <= k*k pushed candidates survive a step;
//   B2 `exclude_num` is per video, global over the search and monotone; a parent expands k - exclude_num children,
//      with exclude_num sampled when that parent's inner loop starts; the search stops once exclude_num >= k
//      (the reference's `== k` test never fires after an overshoot, but then no child is ever generated again);
//   B3 a finished hypothesis is only seen if <eos> is among the parent's first k - exclude_num words;
//   B4 score = logprob / len**lnf for finished hypotheses when lnf > 0 (len counts the <eos>);
//   B5 ordering by score only (ties: first pushed wins here); B6 p = softmax(logits) in fp32, log in fp64;
//   B7 the returned sentence keeps its trailing <eos>.
// Rows are beam-major: row = j*B + b (so the per-video G2x addend is row % B).
#pragma once

#define BEAM_MAXK 8

struct BeamState {
    int B, k, Tc;
    int* nlive;        // [B]
    int* exclude;      // [B]
    int* done;         // [B]
    double* live_lp;   // [k*B]  beam-major
    int* live_len;     // [k*B]
    int* sent[2];      // [k*B, Tc] ping-pong
    int* parent;       // [k*B] parent row (beam index) chosen for the next step
    int* tok;          // [k*B] last word of each live beam (input of the next step)
    // best finished hypothesis
    int* fin_has;      // [B]
    double* fin_score; // [B]
    double* fin_lp;    // [B]
    int* fin_len;      // [B]
    int* fin_sent;     // [B, Tc]
};

// One thread per video: consume the top-k words of each live parent (rows j*B+b, already in descending score
// order), update finals / exclude_num, and select the next k live beams.
// The top-k arrays are indexed [(j * tB + tb) * k + c]: (B, b) for the global [rows, k] arrays, (1, 0) for a per-video copy.
// COPY_SENT = false leaves the children's word histories to the caller (beam_step_kernel copies them with the whole CTA).
template <bool COPY_SENT>
__device__ void beam_update_one(const BeamState& s, const int* top_idx, const float* top_lp, int tB, int tb, int step, int cur, double lnf, int b) {
    const int k = s.k, B = s.B, Tc = s.Tc;
    if (s.done[b]) { s.nlive[b] = 0; return; }
    double c_score[BEAM_MAXK * BEAM_MAXK];
    int c_par[BEAM_MAXK * BEAM_MAXK], c_word[BEAM_MAXK * BEAM_MAXK];
    int nc = 0;
    int excl = s.exclude[b];
    const int nl = s.nlive[b];
    const int* sent_cur = s.sent[cur];
    for (int j = 0; j < nl; ++j) {
        const int row = j * B + b;
        const int nchild = step == 0 ? k : k - excl;          // range(beam_size - exclude_num) evaluated once per parent
        for (int c = 0; c < nchild; ++c) {
            int w = top_idx[(j * tB + tb) * k + c];
            double lp = s.live_lp[row] + (double)top_lp[(j * tB + tb) * k + c];
            int len = s.live_len[row] + 1;
            if (step > 0 && w == 0) {                         // finished hypothesis -> final_captions.push
                double sc = lp;
                if (lnf > 0.0) sc = lp / pow((double)len, lnf);
                if (!s.fin_has[b] || sc > s.fin_score[b]) {
                    s.fin_has[b] = 1; s.fin_score[b] = sc; s.fin_lp[b] = lp; s.fin_len[b] = len;
                    for (int t = 0; t < len - 1; ++t) s.fin_sent[b * Tc + t] = sent_cur[(size_t)row * Tc + t];
                    s.fin_sent[b * Tc + len - 1] = 0;
                    for (int t = len; t < Tc; ++t) s.fin_sent[b * Tc + t] = 0;
                }
                excl += 1;
            } else {                                          // captions.push
                c_score[nc] = lp; c_par[nc] = j; c_word[nc] = w; ++nc;
            }
        }
    }
    s.exclude[b] = excl;
    // mid_captions = captions.extract(sort=True)[:beam_size]: best k by score, earlier push wins ties
    int* sent_nxt = s.sent[cur ^ 1];
    int taken = 0;
    double new_lp[BEAM_MAXK]; int new_len[BEAM_MAXK], new_par[BEAM_MAXK], new_word[BEAM_MAXK];
    for (int j = 0; j < k && j < nc; ++j) {
        int best = -1;
        for (int c = 0; c < nc; ++c)
            if (c_par[c] >= 0 && (best < 0 || c_score[c] > c_score[best])) best = c;
        new_lp[j] = c_score[best]; new_par[j] = c_par[best]; new_word[j] = c_word[best];
        new_len[j] = s.live_len[c_par[best] * B + b] + 1;
        c_par[best] = -1;
        ++taken;
    }
    if (COPY_SENT)
        for (int j = 0; j < taken; ++j) {
            const int row = j * B + b, prow = new_par[j] * B + b;
            for (int t = 0; t < new_len[j] - 1; ++t) sent_nxt[(size_t)row * Tc + t] = sent_cur[(size_t)prow * Tc + t];
            sent_nxt[(size_t)row * Tc + new_len[j] - 1] = new_word[j];
        }
    for (int j = 0; j < k; ++j) {
        const int row = j * B + b;
        if (j < taken) { s.live_lp[row] = new_lp[j]; s.live_len[row] = new_len[j]; s.parent[row] = new_par[j]; s.tok[row] = new_word[j]; }
        else { s.parent[row] = 0; s.tok[row] = 0; }
    }
    s.nlive[b] = taken;
    if (excl >= k) s.done[b] = 1;
}
__global__ void beam_update_kernel(BeamState s, const int* __restrict__ top_idx, const float* __restrict__ top_lp, int step, int cur, double lnf) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < s.B) beam_update_one<true>(s, top_idx, top_lp, s.B, b, step, cur, lnf, b);
}

// Fused beam step (default): one CTA per video.
//   1. warp j merges the per-part candidates of EpiLogitsTopK for live parent j: lse over the parts' (max, sum) pairs, then k
//      rounds of warp arg-max, round q taking the best candidate that comes strictly after round q-1's winner in the
//      (value descending, index ascending) order -- so no "taken" list is needed and ties resolve like tf.nn.top_k;
//   2. thread 0 runs the reference's bookkeeping for the video (beam_update_one) on the shared-memory copy of the top-k lists;
//   3. the whole CTA copies the word histories and the LSTM2 state rows of the chosen parents into the rows of their children
//      (a video's children only ever descend from that video's parents, so no grid-wide dependency exists).
// The next step's cell kernel is a programmatic dependent: it may start its prologue and weight prefetch at once.
static_assert(BEAM_MAXK == TOPK_MAX, "EpiLogitsTopK keeps BEAM_MAXK candidates per part");
template <typename T>
__global__ void __launch_bounds__(32 * BEAM_MAXK) beam_step_kernel(BeamState s, const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                                   const float2* __restrict__ stat, int nparts, int V,
                                                                   int step, int cur, double lnf, const T* __restrict__ h_new, T* __restrict__ h_next,
                                                                   const float* __restrict__ c_new, float* __restrict__ c_next, int Hp) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ int top_idx[BEAM_MAXK * BEAM_MAXK];
    __shared__ float top_lp[BEAM_MAXK * BEAM_MAXK];
    const int b = blockIdx.x, B = s.B, k = s.k, Tc = s.Tc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nl = s.done[b] ? 0 : s.nlive[b];
    if (warp < nl) {
        const int row = warp * B + b;
        const float2* sp = stat + (size_t)row * nparts;
        float mx = -INFINITY;
        for (int q = lane; q < nparts; q += 32) mx = fmaxf(mx, sp[q].x);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int q = lane; q < nparts; q += 32) { const float2 t = sp[q]; se += t.y * expf(t.x - mx); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float lse = mx + logf(se);
        const float* cv = cand_val + (size_t)row * nparts * BEAM_MAXK;
        const int* ci = cand_idx + (size_t)row * nparts * BEAM_MAXK;
        const int nc = nparts * BEAM_MAXK;
        ArgVal prev; prev.v = INFINITY; prev.i = -1;
        for (int j = 0; j < k; ++j) {
            ArgVal best; best.v = -INFINITY; best.i = 0x7fffffff;
            for (int c = lane; c < nc; c += 32) {
                ArgVal x; x.v = cv[c]; x.i = ci[c];
                const bool after = x.v < prev.v || (x.v == prev.v && x.i > prev.i);
                if (after && x.i < V) best = argmax_op(best, x);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgVal t;
                t.v = __shfl_xor_sync(0xffffffffu, best.v, o);
                t.i = __shfl_xor_sync(0xffffffffu, best.i, o);
                best = argmax_op(best, t);
            }
            if (lane == 0) { top_idx[warp * k + j] = best.i; top_lp[warp * k + j] = best.v - lse; }
            prev = best;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) beam_update_one<false>(s, top_idx, top_lp, 1, 0, step, cur, lnf, b);
    __syncthreads();
    {   // word histories of the surviving children: parent's words + the new one
        const int taken = s.nlive[b];
        const int* sent_cur = s.sent[cur];
        int* sent_nxt = s.sent[cur ^ 1];
        for (int u = threadIdx.x; u < taken * Tc; u += blockDim.x) {
            const int j = u / Tc, t = u % Tc, row = j * B + b, len = s.live_len[row];
            if (t < len - 1) sent_nxt[(size_t)row * Tc + t] = sent_cur[(size_t)(s.parent[row] * B + b) * Tc + t];
            else if (t == len - 1) sent_nxt[(size_t)row * Tc + t] = s.tok[row];
        }
    }
    constexpr int HV = 16 / sizeof(T);
    for (int j = 0; j < k; ++j) {
        const int row = j * B + b, prow = s.parent[row] * B + b;
        const uint4* hs = reinterpret_cast<const uint4*>(h_new + (size_t)prow * Hp);
        uint4* hd = reinterpret_cast<uint4*>(h_next + (size_t)row * Hp);
        for (int c = threadIdx.x; c < Hp / HV; c += blockDim.x) hd[c] = hs[c];
        const float4* cs = reinterpret_cast<const float4*>(c_new + (size_t)prow * Hp);
        float4* cd = reinterpret_cast<float4*>(c_next + (size_t)row * Hp);
        for (int c = threadIdx.x; c < Hp / 4; c += blockDim.x) cd[c] = cs[c];
    }
}

// next-step LSTM2 state rows: out[j*B+b] = in[parent[j*B+b]*B + b]
template <typename U>
__global__ void beam_gather_kernel(const U* __restrict__ in, const int* __restrict__ parent, int B, int ld, U* __restrict__ out) {
    int row = blockIdx.x, b = row % B;
    const U* s = in + (size_t)(parent[row] * B + b) * ld;
    U* d = out + (size_t)row * ld;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) d[c] = s[c];
}

__global__ void beam_init_kernel(BeamState s) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    s.nlive[b] = 1; s.exclude[b] = 0; s.done[b] = 0; s.fin_has[b] = 0; s.fin_score[b] = 0.0; s.fin_lp[b] = 0.0; s.fin_len[b] = 0;
    for (int j = 0; j < s.k; ++j) { s.live_lp[j * s.B + b] = 0.0; s.live_len[j * s.B + b] = 0; s.tok[j * s.B + b] = 1; s.parent[j * s.B + b] = 0; }
}

// if not final_captions.size(): final_captions = captions  ->  best surviving candidate of the last step
__global__ void beam_finish_kernel(BeamState s, int cur, int* __restrict__ sentences, int* __restrict__ lengths, float* __restrict__ logprob,
                                   float* __restrict__ score) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    const int Tc = s.Tc;
    if (s.fin_has[b]) {
        for (int t = 0; t < Tc; ++t) sentences[b * Tc + t] = s.fin_sent[b * Tc + t];
        lengths[b] = s.fin_len[b]; logprob[b] = (float)s.fin_lp[b]; score[b] = (float)s.fin_score[b];
    } else {
        int len = s.live_len[b];     // beam 0 = best live candidate
        for (int t = 0; t < Tc; ++t) sentences[b * Tc + t] = t < len ? s.sent[cur][(size_t)b * Tc + t] : 0;
        lengths[b] = len; logprob[b] = (float)s.live_lp[b]; score[b] = (float)s.live_lp[b];
    }
}

template <typename T>
static int beam_impl(s2vt_handle* h, cudaStream_t st, const float* video, int B, int k, float lnf, int32_t* sentences_out, int32_t* lengths_out,
                     float* logprob_out, float* score_out) {
    typedef typename Fwd<T>::type F;   // forward operands: fp16 in the bf16 mode (common.cuh)
    const int Tv = h->Tv, Tc = h->Tc, Hp = h->Hp, Gp = h->Gp, Vp = h->Vp;
    const int R = B * k;
    Arena a(h->ws, h->ws_bytes);
    Roll<T> r;
    plan_roll<T>(h, a, B, R, r);
    BeamState s;
    s.B = B; s.k = k; s.Tc = Tc;
    s.nlive = a.take<int>(B); s.exclude = a.take<int>(B); s.done = a.take<int>(B);
    s.live_lp = a.take<double>(R); s.live_len = a.take<int>(R);
    s.sent[0] = a.take<int>((size_t)R * Tc); s.sent[1] = a.take<int>((size_t)R * Tc);
    s.parent = a.take<int>(R); s.tok = a.take<int>(R);
    s.fin_has = a.take<int>(B); s.fin_score = a.take<double>(B); s.fin_lp = a.take<double>(B); s.fin_len = a.take<int>(B);
    s.fin_sent = a.take<int>((size_t)B * Tc);
    int* top_idx = a.take<int>((size_t)R * k);
    float* top_lp = a.take<float>((size_t)R * k);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small: need %zu bytes", a.used);
    TRY(run_encoder<T>(h, st, video, B, r));
    beam_init_kernel<<<(B + 127) / 128, 128, 0, st>>>(s); KCHECK(h);
    // every beam starts from the encoder state; step 0 only uses beam 0 (nlive = 1)
    tile_rows_kernel<F><<<R, 256, 0, st>>>(r.h2_final, B, R, Hp, r.h2r[0]); KCHECK(h);
    tile_rows_kernel<float><<<R, 256, 0, st>>>(r.c2e[Tv & 1], B, R, Hp, r.c2r[0]); KCHECK(h);
    int cur = 0;   // sentence ping-pong index
    // fused step (default; s2vt_set_overlap bit 3 keeps the un-fused launches as the A/B checker): the candidate arrays live in
    // the logits buffer, which the fused path never writes (nparts * (2 * BEAM_MAXK + 2) floats per row < Vp)
    const bool fused = !(h->overlap & 8);
    const int nparts = Vp / (logits_tile_bn<T>(h) / EpiLogitsTopK<T>::PARTS);
    float* cand_val = r.logits;
    int* cand_idx = reinterpret_cast<int*>(cand_val + (size_t)R * nparts * BEAM_MAXK);
    float2* cand_stat = reinterpret_cast<float2*>(cand_idx + (size_t)R * nparts * BEAM_MAXK);
    for (int i = 0; i < Tc; ++i) {
        const int t = Tv + i;
        const int rows = i == 0 ? B : R;
        typename EpiLstmFwd<F>::Params ep;
        memset(&ep, 0, sizeof ep);
        ep.M = rows; ep.Hp = Hp; ep.bias = h->b2_p; ep.add0 = r.G2x + (size_t)t * B * Gp; ep.add0_mod = B;
        ep.add1 = reinterpret_cast<const F*>(h->Etab); ep.tok = s.tok;
        ep.c_prev = r.c2r[0]; ep.c_out = r.c2r[1]; ep.h_out = r.h2r[1]; ep.keep = 1.f;
        TRY((gemm<F, CfgStep, EpiLstmFwd<F>>(h, st, r.h2r[0], Hp, h->W2hT, Hp, rows, Gp, Hp, ep)));
        if (fused) {
            // :213-217 logits -> softmax -> top_k, fused: per-part candidates + soft-max statistics instead of the [rows, V] logits
            typename EpiLogitsTopK<T>::Params el = {rows, h->V, h->bo_p, cand_val, cand_idx, cand_stat, nparts};
            TRY((gemm<F, CfgBig, EpiLogitsTopK<T>>(h, st, r.h2r[1], Hp, h->WoT, Hp, rows, Vp, Hp, el)));
            beam_step_kernel<F><<<B, 32 * BEAM_MAXK, 0, st>>>(s, cand_val, cand_idx, cand_stat, nparts, h->V, i, cur, (double)lnf,
                                                             r.h2r[1], r.h2r[0], r.c2r[1], r.c2r[0], Hp); KCHECK(h);
            h->launches++;
            cur ^= 1;
            continue;
        }
        typename EpiStore<F>::Params el = {r.logits, nullptr, Vp, h->bo_p, rows, 0};
        TRY((gemm<F, CfgBig, EpiStore<F>>(h, st, r.h2r[1], Hp, h->WoT, Hp, rows, Vp, Hp, el)));
        topk_rows_kernel<<<rows, ROW_THREADS, 0, st>>>(r.logits, Vp, h->V, k, top_idx, top_lp); KCHECK(h);
        beam_update_kernel<<<(B + 63) / 64, 64, 0, st>>>(s, top_idx, top_lp, i, cur, (double)lnf); KCHECK(h);
        cur ^= 1;
        // children inherit the state produced by their parent's step
        beam_gather_kernel<F><<<R, 256, 0, st>>>(r.h2r[1], s.parent, B, Hp, r.h2r[0]); KCHECK(h);
        beam_gather_kernel<float><<<R, 256, 0, st>>>(r.c2r[1], s.parent, B, Hp, r.c2r[0]); KCHECK(h);
    }
    beam_finish_kernel<<<(B + 127) / 128, 128, 0, st>>>(s, cur, sentences_out, lengths_out, logprob_out, score_out); KCHECK(h);
    return 0;
}

extern "C" int s2vt_beam_search(s2vt_handle* h, const float* video, int B, int beam_size, float lnf, int32_t* sentences_out, int32_t* lengths_out,
                                float* logprob_out, float* score_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!video || B <= 0 || beam_size < 1 || beam_size > BEAM_MAXK || !sentences_out || !lengths_out || !logprob_out || !score_out)
        return h->fail(S2VT_EINVAL, "bad beam_search arguments (1 <= beam_size <= %d)", BEAM_MAXK);
    return DISPATCH(h, beam_impl<bf16>(h, (cudaStream_t)st, video, B, beam_size, lnf, sentences_out, lengths_out, logprob_out, score_out),
                    beam_impl<float>(h, (cudaStream_t)st, video, B, beam_size, lnf, sentences_out, lengths_out, logprob_out, score_out));
}

// ---- single-hypothesis drop-in contracts (batch 1, as the reference runs them) -------------------------------------
// state layout [1, 2H] = concat(c, h) (state_is_tuple=False, tf_s2vt.py:74)
template <typename T>
__global__ void state_unpack_kernel(const float* __restrict__ state, int H, int Hp, float* __restrict__ c, T* __restrict__ hh) {
    for (int u = threadIdx.x; u < Hp; u += blockDim.x) {
        c[u] = u < H ? state[u] : 0.f;
        hh[u] = from_f32<T>(u < H ? state[H + u] : 0.f);
    }
}
__global__ void state_pack_kernel(const float* __restrict__ c, const float* __restrict__ hF, int H, float* __restrict__ state) {
    for (int u = threadIdx.x; u < H; u += blockDim.x) { state[u] = c[u]; state[H + u] = hF[u]; }
}
__global__ void exp_kernel(const float* __restrict__ lp, int n, float* __restrict__ p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = expf(lp[i]);
}
template <typename T>
__global__ void to_f32_kernel(const T* __restrict__ in, int n, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = to_f32(in[i]);
}

template <typename T>
static int beam_init_impl(s2vt_handle* h, cudaStream_t st, const float* video, float* state1_out, float* state2_out) {
    typedef typename Fwd<T>::type F;   // forward operands: fp16 in the bf16 mode (common.cuh)
    const int Tv = h->Tv, Hp = h->Hp, H = h->H;
    Arena a(h->ws, h->ws_bytes);
    Roll<T> r;
    plan_roll<T>(h, a, 1, 1, r);
    float* hF = a.take<float>(Hp);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small");
    TRY(run_encoder<T>(h, st, video, 1, r));
    to_f32_kernel<F><<<(Hp + 255) / 256, 256, 0, st>>>(r.f.h1_all + (size_t)Tv * Hp, Hp, hF); KCHECK(h);
    state_pack_kernel<<<1, 256, 0, st>>>(r.f.c1_all + (size_t)Tv * Hp, hF, H, state1_out); KCHECK(h);
    to_f32_kernel<F><<<(Hp + 255) / 256, 256, 0, st>>>(r.h2_final, Hp, hF); KCHECK(h);
    state_pack_kernel<<<1, 256, 0, st>>>(r.c2e[Tv & 1], hF, H, state2_out); KCHECK(h);
    return 0;
}

template <typename T>
static int beam_step_impl(s2vt_handle* h, cudaStream_t st, const float* state2, const float* state1, const int32_t* word, int k, int32_t* idx_out,
                          float* prob_out, float* state2_out, float* state1_out) {
    typedef typename Fwd<T>::type F;   // forward operands: fp16 in the bf16 mode (common.cuh)
    const int Hp = h->Hp, H = h->H, Gp = h->Gp, Vp = h->Vp;
    h->front_valid = false;
    TRY(wait_late_weights(h, st));
    Arena a(h->ws, h->ws_bytes);
    float* c1 = a.take<float>(Hp); F* h1 = a.take<F>(Hp); float* c1n = a.take<float>(Hp); F* h1n = a.take<F>(Hp); float* h1F = a.take<float>(Hp);
    float* c2 = a.take<float>(Hp); F* h2 = a.take<F>(Hp); float* c2n = a.take<float>(Hp); F* h2n = a.take<F>(Hp); float* h2F = a.take<float>(Hp);
    float* g2x = a.take<float>(Gp); float* logits = a.take<float>(Vp); float* lp = a.take<float>(BEAM_MAXK);
    if (a.overflow) return h->fail(S2VT_ENOSPACE, "workspace too small");
    state_unpack_kernel<F><<<1, 256, 0, st>>>(state1, H, Hp, c1, h1); KCHECK(h);
    state_unpack_kernel<F><<<1, 256, 0, st>>>(state2, H, Hp, c2, h2); KCHECK(h);
    typename EpiLstmFwd<F>::Params e1;
    memset(&e1, 0, sizeof e1);
    e1.M = 1; e1.Hp = Hp; e1.bias = h->b1_p; e1.c_prev = c1; e1.c_out = c1n; e1.h_out = h1n; e1.h_outF = h1F; e1.keep = 1.f;
    TRY((gemm<F, CfgStep, EpiLstmFwd<F>>(h, st, h1, Hp, h->W1hT, Hp, 1, Gp, Hp, e1)));       // lstm1(padding, state1)
    typename EpiStore<F>::Params eg = {g2x, nullptr, Gp, nullptr, 1, 0};
    TRY((gemm<F, CfgStep, EpiStore<F>>(h, st, h1n, Hp, h->W2xT, Hp, 1, Gp, Hp, eg)));
    typename EpiLstmFwd<F>::Params e2;
    memset(&e2, 0, sizeof e2);
    e2.M = 1; e2.Hp = Hp; e2.bias = h->b2_p; e2.add0 = g2x; e2.add1 = reinterpret_cast<const F*>(h->Etab); e2.tok = word;
    e2.c_prev = c2; e2.c_out = c2n; e2.h_out = h2n; e2.h_outF = h2F; e2.keep = 1.f;
    TRY((gemm<F, CfgStep, EpiLstmFwd<F>>(h, st, h2, Hp, h->W2hT, Hp, 1, Gp, Hp, e2)));       // lstm2([out1, emb], state2)
    typename EpiStore<F>::Params el = {logits, nullptr, Vp, h->bo_p, 1, 0};
    TRY((gemm<F, CfgStep, EpiStore<F>>(h, st, h2n, Hp, h->WoT, Hp, 1, Vp, Hp, el)));
    topk_rows_kernel<<<1, ROW_THREADS, 0, st>>>(logits, Vp, h->V, k, idx_out, lp); KCHECK(h);
    exp_kernel<<<1, 32, 0, st>>>(lp, k, prob_out); KCHECK(h);
    state_pack_kernel<<<1, 256, 0, st>>>(c1n, h1F, H, state1_out); KCHECK(h);
    state_pack_kernel<<<1, 256, 0, st>>>(c2n, h2F, H, state2_out); KCHECK(h);
    return 0;
}

extern "C" int s2vt_beam_init(s2vt_handle* h, const float* video, float* state1_out, float* state2_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!video || !state1_out || !state2_out) return h->fail(S2VT_EINVAL, "bad beam_init arguments");
    return DISPATCH(h, beam_init_impl<bf16>(h, (cudaStream_t)st, video, state1_out, state2_out),
                    beam_init_impl<float>(h, (cudaStream_t)st, video, state1_out, state2_out));
}
extern "C" int s2vt_beam_step(s2vt_handle* h, const float* state2, const float* state1, const int32_t* word, int beam_size, int32_t* idx_out,
                              float* prob_out, float* state2_out, float* state1_out, s2vt_stream st) {
    TRY(check_ready(h));
    if (!state2 || !state1 || !word || beam_size < 1 || beam_size > BEAM_MAXK || !idx_out || !prob_out || !state2_out || !state1_out)
        return h->fail(S2VT_EINVAL, "bad beam_step arguments");
    return DISPATCH(h, beam_step_impl<bf16>(h, (cudaStream_t)st, state2, state1, word, beam_size, idx_out, prob_out, state2_out, state1_out),
                    beam_step_impl<float>(h, (cudaStream_t)st, state2, state1, word, beam_size, idx_out, prob_out, state2_out, state1_out));
}
