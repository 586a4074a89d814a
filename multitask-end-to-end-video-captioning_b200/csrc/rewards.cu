// GPU BLEU-4 and ROUGE-L rewards: replace Bleu(4).compute_score / Rouge().compute_score behind the
// evaluate_captions_cider (sic) of bleu_evaluation.py:60-87 and rouge_evaluation.py:60-87 (third-party pycocoevalcap,
// algorithm restated in oracle/bleu_rouge.py).  Same conventions as ciderd.cu: token ids, exact 64-bit n-gram keys, the
// per-video reference tables built once on the host and kept in HBM, one CTA per hypothesis, fp64 arithmetic.
//   BLEU : per video the clipped-count table {n-gram -> max count over its references} (sorted keys) + the reference lengths
//   ROUGE: per video the reference token sequences; LCS by the bit-parallel recurrence (hypothesis <= 64 tokens = one word)
#include <math.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "engine.cuh"

#define RW_N 4
#define RW_MAXTOK 64
#define RW_MAXG (4 * RW_MAXTOK)

struct RewardHeader {
    int64_t n_videos, n_refs, n_tokens, n_entries;
    int64_t off_video_ref, off_ref_tok, off_tokens, off_video_entry, off_ent_keys, off_ent_max, total_bytes;
};

struct s2vt_reward_corpus {
    RewardHeader hd;
    std::vector<int64_t> video_ref, ref_tok, video_entry;
    std::vector<int32_t> tokens, ent_max;
    std::vector<unsigned long long> ent_keys;
};

extern "C" int s2vt_reward_corpus_create(const int32_t* ref_tokens, const int64_t* ref_offsets, int64_t n_refs, const int64_t* video_ref_offsets,
                                         int64_t n_videos, s2vt_reward_corpus** out) {
    if (!ref_tokens || !ref_offsets || !video_ref_offsets || !out || n_refs <= 0 || n_videos <= 0) return S2VT_EINVAL;
    for (int64_t i = 0; i < ref_offsets[n_refs]; ++i)
        if (ref_tokens[i] < 0 || ref_tokens[i] >= 65535) return S2VT_EINVAL;
    s2vt_reward_corpus* c = new s2vt_reward_corpus();
    c->video_ref.assign(video_ref_offsets, video_ref_offsets + n_videos + 1);
    c->ref_tok.assign(ref_offsets, ref_offsets + n_refs + 1);
    c->tokens.assign(ref_tokens, ref_tokens + ref_offsets[n_refs]);
    c->video_entry.resize(n_videos + 1);
    for (int64_t v = 0; v < n_videos; ++v) {      // cook_refs: maxcounts[ngram] = max over the references of the video
        c->video_entry[v] = (int64_t)c->ent_keys.size();
        std::unordered_map<unsigned long long, int> maxc;
        for (int64_t r = video_ref_offsets[v]; r < video_ref_offsets[v + 1]; ++r) {
            std::unordered_map<unsigned long long, int> cnt;
            const int32_t* tok = ref_tokens + ref_offsets[r];
            const int64_t len = ref_offsets[r + 1] - ref_offsets[r];
            for (int n = 1; n <= RW_N; ++n)
                for (int64_t i = 0; i + n <= len; ++i) {
                    unsigned long long key = 0;
                    for (int j = 0; j < n; ++j) key |= (unsigned long long)(tok[i + j] + 1) << (16 * j);
                    cnt[key] += 1;
                }
            for (auto& kv : cnt) { int& m = maxc[kv.first]; if (kv.second > m) m = kv.second; }
        }
        std::vector<std::pair<unsigned long long, int>> e(maxc.begin(), maxc.end());
        std::sort(e.begin(), e.end());
        for (auto& kv : e) { c->ent_keys.push_back(kv.first); c->ent_max.push_back(kv.second); }
    }
    c->video_entry[n_videos] = (int64_t)c->ent_keys.size();
    RewardHeader& h = c->hd;
    h.n_videos = n_videos; h.n_refs = n_refs; h.n_tokens = (int64_t)c->tokens.size(); h.n_entries = (int64_t)c->ent_keys.size();
    size_t o = ru64(sizeof(RewardHeader), 256);
    auto place = [&](size_t bytes) { size_t at = o; o += ru64(bytes, 256); return (int64_t)at; };
    h.off_video_ref = place((n_videos + 1) * 8); h.off_ref_tok = place((n_refs + 1) * 8); h.off_tokens = place(h.n_tokens * 4);
    h.off_video_entry = place((n_videos + 1) * 8); h.off_ent_keys = place(h.n_entries * 8); h.off_ent_max = place(h.n_entries * 4);
    h.total_bytes = (int64_t)o;
    *out = c;
    return S2VT_OK;
}
extern "C" void s2vt_reward_corpus_destroy(s2vt_reward_corpus* c) { delete c; }
extern "C" size_t s2vt_reward_corpus_device_bytes(const s2vt_reward_corpus* c) { return c ? (size_t)c->hd.total_bytes : 0; }
extern "C" int s2vt_reward_corpus_serialize(const s2vt_reward_corpus* c, void* host_buffer) {
    if (!c || !host_buffer) return S2VT_EINVAL;
    char* b = (char*)host_buffer;
    memset(b, 0, (size_t)c->hd.total_bytes);
    memcpy(b, &c->hd, sizeof(RewardHeader));
    memcpy(b + c->hd.off_video_ref, c->video_ref.data(), c->video_ref.size() * 8);
    memcpy(b + c->hd.off_ref_tok, c->ref_tok.data(), c->ref_tok.size() * 8);
    memcpy(b + c->hd.off_tokens, c->tokens.data(), c->tokens.size() * 4);
    memcpy(b + c->hd.off_video_entry, c->video_entry.data(), c->video_entry.size() * 8);
    memcpy(b + c->hd.off_ent_keys, c->ent_keys.data(), c->ent_keys.size() * 8);
    memcpy(b + c->hd.off_ent_max, c->ent_max.data(), c->ent_max.size() * 4);
    return S2VT_OK;
}

// ---- BLEU: BleuScorer.compute_score(option='closest') per-sentence list ----------------------------------------------
__global__ void __launch_bounds__(128) bleu_score_kernel(const char* __restrict__ corpus, const int* __restrict__ hyp, const int* __restrict__ video_of_row,
                                                         int Tc, double* __restrict__ bleu_out, int* __restrict__ comps_out) {
    const RewardHeader* hd = reinterpret_cast<const RewardHeader*>(corpus);
    const long long* video_ref = reinterpret_cast<const long long*>(corpus + hd->off_video_ref);
    const long long* ref_tok = reinterpret_cast<const long long*>(corpus + hd->off_ref_tok);
    const long long* video_entry = reinterpret_cast<const long long*>(corpus + hd->off_video_entry);
    const unsigned long long* ent_keys = reinterpret_cast<const unsigned long long*>(corpus + hd->off_ent_keys);
    const int* ent_max = reinterpret_cast<const int*>(corpus + hd->off_ent_max);

    __shared__ int tok[RW_MAXTOK];
    __shared__ unsigned long long keys[RW_MAXG];
    __shared__ int L, G;
    __shared__ int correct[RW_N];
    const int row = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {   // words before the first <eos> (decode_captions, cider_evaluation.py:122-143)
        int l = 0;
        while (l < Tc && l < RW_MAXTOK && hyp[(size_t)row * Tc + l] != 0) { tok[l] = hyp[(size_t)row * Tc + l]; ++l; }
        L = l;
        int g = 0;
        for (int n = 1; n <= RW_N; ++n) g += l >= n ? l - n + 1 : 0;
        G = g;
        for (int n = 0; n < RW_N; ++n) correct[n] = 0;
    }
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        int n = 1, rem = g;
        while (rem >= L - n + 1) { rem -= L - n + 1; ++n; }
        unsigned long long key = 0;
        for (int j = 0; j < n; ++j) key |= (unsigned long long)(tok[rem + j] + 1) << (16 * j);
        keys[g] = key;
    }
    __syncthreads();
    const int vid = video_of_row[row];
    const long long e0 = video_entry[vid], e1 = video_entry[vid + 1];
    // cook_test: correct[k] += min(refmaxcounts.get(ngram, 0), count) over the distinct n-grams (first occurrence owns it)
    for (int g = tid; g < G; g += blockDim.x) {
        const unsigned long long key = keys[g];
        int cnt = 0; bool first = true;
        for (int j = 0; j < G; ++j)
            if (keys[j] == key) { ++cnt; if (j < g) first = false; }
        if (!first) continue;
        long long lo = e0, hi = e1; int mx = 0;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            unsigned long long k = ent_keys[mid];
            if (k == key) { mx = ent_max[mid]; break; }
            if (k < key) lo = mid + 1; else hi = mid;
        }
        const int n = key >> 48 ? 3 : key >> 32 ? 2 : key >> 16 ? 1 : 0;
        atomicAdd(&correct[n], cnt < mx ? cnt : mx);
    }
    __syncthreads();
    if (tid == 0) {
        const int testlen = L;
        long long best_d = -1; long long reflen = 0;        // min((abs(l - testlen), l) for l in reflens)[1]
        for (long long r = video_ref[vid]; r < video_ref[vid + 1]; ++r) {
            long long l = ref_tok[r + 1] - ref_tok[r];
            long long d = l > testlen ? l - testlen : testlen - l;
            if (best_d < 0 || d < best_d || (d == best_d && l < reflen)) { best_d = d; reflen = l; }
        }
        const double tiny = 1e-15, small = 1e-9;
        const double ratio = ((double)testlen + tiny) / ((double)reflen + small);
        const double bp = ratio < 1.0 ? exp(1.0 - 1.0 / ratio) : 1.0;
        if (comps_out) {   // cook_test components: the corpus-level score sums them over the sentences (BleuScorer.compute_score totalcomps)
            int* co = comps_out + (size_t)row * 10;
            for (int k = 0; k < RW_N; ++k) { co[k] = correct[k]; co[4 + k] = testlen - k > 0 ? testlen - k : 0; }
            co[8] = testlen; co[9] = (int)reflen;
        }
        double bleu = 1.0;
        for (int k = 0; k < RW_N; ++k) {
            const int guess = testlen - k > 0 ? testlen - k : 0;
            bleu *= ((double)correct[k] + tiny) / ((double)guess + small);
            double v = k == 0 ? bleu : pow(bleu, 1.0 / (double)(k + 1));
            bleu_out[(size_t)row * RW_N + k] = ratio < 1.0 ? v * bp : v;
        }
    }
}

// ---- ROUGE-L: Rouge.calc_score ------------------------------------------------------------------------------------------
// One warp per (hypothesis, reference) pair, lanes over hypothesis positions: M = ballot(hyp[pos] == ref token) gives the
// match word of the Allison-Dix / Hyyro bit-vector LCS:  U = V & M;  V = (V + U) | (V - U);  LCS = #zero bits of V in [0, L).
__global__ void __launch_bounds__(128) rouge_score_kernel(const char* __restrict__ corpus, const int* __restrict__ hyp, const int* __restrict__ video_of_row,
                                                          int Tc, int empty_token, double* __restrict__ out) {
    const RewardHeader* hd = reinterpret_cast<const RewardHeader*>(corpus);
    const long long* video_ref = reinterpret_cast<const long long*>(corpus + hd->off_video_ref);
    const long long* ref_tok = reinterpret_cast<const long long*>(corpus + hd->off_ref_tok);
    const int* tokens = reinterpret_cast<const int*>(corpus + hd->off_tokens);
    __shared__ int tok[RW_MAXTOK];
    __shared__ int L;
    __shared__ unsigned long long best_p, best_r;      // max over references of lcs/len_c and lcs/len_r, as fp64 bit patterns (non-negative)
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) {
        int l = 0;
        while (l < Tc && l < RW_MAXTOK && hyp[(size_t)row * Tc + l] != 0) { tok[l] = hyp[(size_t)row * Tc + l]; ++l; }
        if (l == 0) { tok[0] = empty_token; l = 1; }     // ''.split(" ") == [''] : one empty token
        L = l; best_p = 0ull; best_r = 0ull;
    }
    __syncthreads();
    const int len_c = L;
    const int t_lo = lane < len_c ? tok[lane] : -1, t_hi = lane + 32 < len_c ? tok[lane + 32] : -1;
    const unsigned long long lmask = len_c >= 64 ? ~0ull : ((1ull << len_c) - 1ull);
    const int vid = video_of_row[row];
    for (long long r = video_ref[vid] + warp; r < video_ref[vid + 1]; r += nwarps) {
        const long long b = ref_tok[r], e = ref_tok[r + 1];
        unsigned long long V = ~0ull;
        for (long long i = b; i < e; ++i) {
            const int c = tokens[i];
            const unsigned long long M = (unsigned long long)__ballot_sync(0xffffffffu, t_lo == c) | ((unsigned long long)__ballot_sync(0xffffffffu, t_hi == c) << 32);
            const unsigned long long U = V & M;
            V = (V + U) | (V - U);
        }
        if (lane == 0) {
            const int lcs = __popcll(~V & lmask);
            const double p = (double)lcs / (double)len_c, rc = (double)lcs / (double)(e - b);
            atomicMax(&best_p, (unsigned long long)__double_as_longlong(p));
            atomicMax(&best_r, (unsigned long long)__double_as_longlong(rc));
        }
    }
    __syncthreads();
    if (tid == 0) {
        const double p = __longlong_as_double((long long)best_p), rc = __longlong_as_double((long long)best_r);
        const double beta2 = 1.2 * 1.2;
        out[row] = (p != 0.0 && rc != 0.0) ? ((1.0 + beta2) * p * rc) / (rc + beta2 * p) : 0.0;
    }
}

extern "C" int s2vt_bleu_score(const void* corpus_device, const int32_t* hyp, const int32_t* video_of_row, int N, int Tc, double* bleu_out, int32_t* comps_out,
                               s2vt_stream st) {
    if (!corpus_device || !hyp || !video_of_row || !bleu_out || N <= 0 || Tc <= 0 || Tc > RW_MAXTOK) return S2VT_EINVAL;
    bleu_score_kernel<<<N, 128, 0, (cudaStream_t)st>>>((const char*)corpus_device, hyp, video_of_row, Tc, bleu_out, comps_out);
    return cudaGetLastError() == cudaSuccess ? S2VT_OK : S2VT_ECUDA;
}

extern "C" int s2vt_rouge_score(const void* corpus_device, const int32_t* hyp, const int32_t* video_of_row, int N, int Tc, int32_t empty_token,
                                double* rouge_out, s2vt_stream st) {
    if (!corpus_device || !hyp || !video_of_row || !rouge_out || N <= 0 || Tc <= 0 || Tc > RW_MAXTOK) return S2VT_EINVAL;
    rouge_score_kernel<<<N, 128, 0, (cudaStream_t)st>>>((const char*)corpus_device, hyp, video_of_row, Tc, empty_token, rouge_out);
    return cudaGetLastError() == cudaSuccess ? S2VT_OK : S2VT_ECUDA;
}
