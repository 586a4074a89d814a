// GPU CIDEr-D reward: replaces CiderD_scorer.compute_score behind evaluate_captions_cider
// (cider_evaluation.py:33-39, 60-87; third-party pyciderevalcap, algorithm restated in oracle/ciderd.py).
//
// n-grams are EXACT 64-bit keys: key = sum_i (token_i + 1) << (16 i), i < n <= 4 (token ids < 65535), so the
// order n is the index of the highest non-zero 16-bit field and "hashed" counts are bit-exact by construction.
// Host (once per corpus): per-reference sorted (key, tf-idf weight) tables, norms, bigram lengths, and the sorted
// document-frequency table.  Device: one CTA per hypothesis; fp64 arithmetic throughout (HBM-bound table scans).
#include <math.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "engine.cuh"

#define CIDER_N 4
#define CIDER_SIGMA 6.0
#define CIDER_MAXTOK 64
#define CIDER_MAXG (4 * CIDER_MAXTOK)

struct CorpusHeader {
    int64_t n_videos, n_refs, n_df, n_entries;
    double ref_len;  // ln(n_videos)
    int64_t off_df_keys, off_df_idf, off_video_ref, off_ref_entry, off_ent_keys, off_ent_vec, off_ref_norm, off_ref_length, total_bytes;
};

struct ciderd_corpus {
    CorpusHeader hd;
    std::vector<unsigned long long> df_keys, ent_keys;
    std::vector<double> df_idf, ent_vec, ref_norm;
    std::vector<int64_t> video_ref, ref_entry;
    std::vector<int32_t> ref_length;
};

static void ngram_counts(const int32_t* tok, int64_t len, std::unordered_map<unsigned long long, int>& counts) {
    for (int n = 1; n <= CIDER_N; ++n)
        for (int64_t i = 0; i + n <= len; ++i) {
            unsigned long long key = 0;
            for (int j = 0; j < n; ++j) key |= (unsigned long long)(tok[i + j] + 1) << (16 * j);
            counts[key] += 1;
        }
}
static inline int key_order(unsigned long long key) { return key >> 48 ? 3 : key >> 32 ? 2 : key >> 16 ? 1 : 0; }

extern "C" int ciderd_corpus_create(const int32_t* ref_tokens, const int64_t* ref_offsets, int64_t n_refs, const int64_t* video_ref_offsets,
                                    int64_t n_videos, const uint8_t* video_in_df, ciderd_corpus** out) {
    if (!ref_tokens || !ref_offsets || !video_ref_offsets || !out || n_refs <= 0 || n_videos <= 0) return S2VT_EINVAL;
    for (int64_t i = 0; i < ref_offsets[n_refs]; ++i)
        if (ref_tokens[i] < 0 || ref_tokens[i] >= 65535) return S2VT_EINVAL;
    ciderd_corpus* c = new ciderd_corpus();
    std::vector<std::unordered_map<unsigned long long, int>> cooked(n_refs);
    for (int64_t r = 0; r < n_refs; ++r) ngram_counts(ref_tokens + ref_offsets[r], ref_offsets[r + 1] - ref_offsets[r], cooked[r]);
    // document frequency: number of videos whose reference set contains the n-gram
    std::unordered_map<unsigned long long, int> df;
    int64_t n_df_videos = 0;
    for (int64_t v = 0; v < n_videos; ++v) {
        if (video_in_df && !video_in_df[v]) continue;
        ++n_df_videos;
        std::unordered_map<unsigned long long, char> seen;
        for (int64_t r = video_ref_offsets[v]; r < video_ref_offsets[v + 1]; ++r)
            for (auto& kv : cooked[r]) seen[kv.first] = 1;
        for (auto& kv : seen) df[kv.first] += 1;
    }
    if (n_df_videos <= 0) { delete c; return S2VT_EINVAL; }
    const double ref_len = log((double)n_df_videos);
    std::vector<std::pair<unsigned long long, int>> dfv(df.begin(), df.end());
    std::sort(dfv.begin(), dfv.end());
    c->df_keys.resize(dfv.size()); c->df_idf.resize(dfv.size());
    for (size_t i = 0; i < dfv.size(); ++i) {
        c->df_keys[i] = dfv[i].first;
        c->df_idf[i] = ref_len - log(std::max(1.0, (double)dfv[i].second));
    }
    c->video_ref.assign(video_ref_offsets, video_ref_offsets + n_videos + 1);
    c->ref_entry.resize(n_refs + 1); c->ref_norm.assign(n_refs * CIDER_N, 0.0); c->ref_length.assign(n_refs, 0);
    for (int64_t r = 0; r < n_refs; ++r) {
        c->ref_entry[r] = (int64_t)c->ent_keys.size();
        std::vector<std::pair<unsigned long long, int>> e(cooked[r].begin(), cooked[r].end());
        std::sort(e.begin(), e.end());
        for (auto& kv : e) {
            auto dit = df.find(kv.first);
            double w = (double)kv.second * (ref_len - log(std::max(1.0, dit == df.end() ? 0.0 : (double)dit->second)));
            int n = key_order(kv.first);
            c->ent_keys.push_back(kv.first); c->ent_vec.push_back(w);
            c->ref_norm[r * CIDER_N + n] += w * w;
            if (n == 1) c->ref_length[r] += kv.second;   // the library's length counts bigrams only
        }
        for (int n = 0; n < CIDER_N; ++n) c->ref_norm[r * CIDER_N + n] = sqrt(c->ref_norm[r * CIDER_N + n]);
    }
    c->ref_entry[n_refs] = (int64_t)c->ent_keys.size();
    CorpusHeader& h = c->hd;
    h.n_videos = n_videos; h.n_refs = n_refs; h.n_df = (int64_t)c->df_keys.size(); h.n_entries = (int64_t)c->ent_keys.size(); h.ref_len = ref_len;
    size_t o = ru64(sizeof(CorpusHeader), 256);
    auto place = [&](size_t bytes) { size_t at = o; o += ru64(bytes, 256); return (int64_t)at; };
    h.off_df_keys = place(h.n_df * 8); h.off_df_idf = place(h.n_df * 8); h.off_video_ref = place((n_videos + 1) * 8);
    h.off_ref_entry = place((n_refs + 1) * 8); h.off_ent_keys = place(h.n_entries * 8); h.off_ent_vec = place(h.n_entries * 8);
    h.off_ref_norm = place(n_refs * CIDER_N * 8); h.off_ref_length = place(n_refs * 4);
    h.total_bytes = (int64_t)o;
    *out = c;
    return S2VT_OK;
}
extern "C" void ciderd_corpus_destroy(ciderd_corpus* c) { delete c; }
extern "C" size_t ciderd_corpus_device_bytes(const ciderd_corpus* c) { return c ? (size_t)c->hd.total_bytes : 0; }
extern "C" int ciderd_corpus_serialize(const ciderd_corpus* c, void* host_buffer) {
    if (!c || !host_buffer) return S2VT_EINVAL;
    char* b = (char*)host_buffer;
    memset(b, 0, (size_t)c->hd.total_bytes);
    memcpy(b, &c->hd, sizeof(CorpusHeader));
    memcpy(b + c->hd.off_df_keys, c->df_keys.data(), c->df_keys.size() * 8);
    memcpy(b + c->hd.off_df_idf, c->df_idf.data(), c->df_idf.size() * 8);
    memcpy(b + c->hd.off_video_ref, c->video_ref.data(), c->video_ref.size() * 8);
    memcpy(b + c->hd.off_ref_entry, c->ref_entry.data(), c->ref_entry.size() * 8);
    memcpy(b + c->hd.off_ent_keys, c->ent_keys.data(), c->ent_keys.size() * 8);
    memcpy(b + c->hd.off_ent_vec, c->ent_vec.data(), c->ent_vec.size() * 8);
    memcpy(b + c->hd.off_ref_norm, c->ref_norm.data(), c->ref_norm.size() * 8);
    memcpy(b + c->hd.off_ref_length, c->ref_length.data(), c->ref_length.size() * 4);
    return S2VT_OK;
}

// index of `key` in sorted keys[lo, hi) or -1
__device__ __forceinline__ long long find_key(const unsigned long long* __restrict__ keys, long long lo, long long hi, unsigned long long key) {
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        unsigned long long k = keys[mid];
        if (k == key) return mid;
        if (k < key) lo = mid + 1; else hi = mid;
    }
    return -1;
}

__global__ void __launch_bounds__(128) ciderd_score_kernel(const char* __restrict__ corpus, const int* __restrict__ hyp, const int* __restrict__ video_of_row,
                                                           int Tc, double* __restrict__ scores, unsigned long long* __restrict__ counts_out) {
    const CorpusHeader* hd = reinterpret_cast<const CorpusHeader*>(corpus);
    const unsigned long long* df_keys = reinterpret_cast<const unsigned long long*>(corpus + hd->off_df_keys);
    const double* df_idf = reinterpret_cast<const double*>(corpus + hd->off_df_idf);
    const long long* video_ref = reinterpret_cast<const long long*>(corpus + hd->off_video_ref);
    const long long* ref_entry = reinterpret_cast<const long long*>(corpus + hd->off_ref_entry);
    const unsigned long long* ent_keys = reinterpret_cast<const unsigned long long*>(corpus + hd->off_ent_keys);
    const double* ent_vec = reinterpret_cast<const double*>(corpus + hd->off_ent_vec);
    const double* ref_norm = reinterpret_cast<const double*>(corpus + hd->off_ref_norm);
    const int* ref_length = reinterpret_cast<const int*>(corpus + hd->off_ref_length);

    __shared__ int tok[CIDER_MAXTOK];
    __shared__ unsigned long long keys[CIDER_MAXG];
    __shared__ unsigned long long ukey[CIDER_MAXG];
    __shared__ double uvec[CIDER_MAXG];
    __shared__ int ucnt[CIDER_MAXG];
    __shared__ int L, G, U, len_h;
    __shared__ double norm_h[CIDER_N], score[CIDER_N];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    if (tid == 0) {   // tokens before the first <eos> (R2)
        int l = 0;
        while (l < Tc && l < CIDER_MAXTOK && hyp[(size_t)row * Tc + l] != 0) { tok[l] = hyp[(size_t)row * Tc + l]; ++l; }
        L = l;
        int g = 0;
        for (int n = 1; n <= CIDER_N; ++n) g += l >= n ? l - n + 1 : 0;
        G = g; U = 0; len_h = l >= 2 ? l - 1 : 0;
        for (int n = 0; n < CIDER_N; ++n) { norm_h[n] = 0.0; score[n] = 0.0; }
    }
    __syncthreads();
    // all n-gram occurrences, ordered by (n, position)
    for (int g = tid; g < G; g += blockDim.x) {
        int n = 1, rem = g;
        while (rem >= L - n + 1) { rem -= L - n + 1; ++n; }
        unsigned long long key = 0;
        for (int j = 0; j < n; ++j) key |= (unsigned long long)(tok[rem + j] + 1) << (16 * j);
        keys[g] = key;
    }
    __syncthreads();
    // multiset -> (unique key, count); first occurrence owns the entry
    for (int g = tid; g < G; g += blockDim.x) {
        unsigned long long key = keys[g];
        int cnt = 0; bool first = true;
        for (int j = 0; j < G; ++j) {
            if (keys[j] == key) { ++cnt; if (j < g) first = false; }
        }
        if (first) {
            int slot = atomicAdd(&U, 1);
            long long at = find_key(df_keys, 0, hd->n_df, key);
            double idf = at >= 0 ? df_idf[at] : hd->ref_len;      // unseen n-gram: df = 0 -> ln(max(1, 0)) = 0
            double w = (double)cnt * idf;
            ukey[slot] = key; uvec[slot] = w; ucnt[slot] = cnt;
            int n = key >> 48 ? 3 : key >> 32 ? 2 : key >> 16 ? 1 : 0;
            atomicAdd(&norm_h[n], w * w);
        }
    }
    __syncthreads();
    if (counts_out) {
        unsigned long long* co = counts_out + (size_t)row * (4 * Tc) * 2;
        for (int g = tid; g < 4 * Tc; g += blockDim.x) {
            co[2 * g] = g < U ? ukey[g] : 0ull;
            co[2 * g + 1] = g < U ? (unsigned long long)ucnt[g] : 0ull;
        }
    }
    const int vid = video_of_row[row];
    const long long r0 = video_ref[vid], r1 = video_ref[vid + 1];
    for (long long r = r0 + warp; r < r1; r += nwarps) {
        double val[CIDER_N] = {0.0, 0.0, 0.0, 0.0};
        const long long e0 = ref_entry[r], e1 = ref_entry[r + 1];
        for (int g = lane; g < U; g += 32) {
            unsigned long long key = ukey[g];
            long long at = find_key(ent_keys, e0, e1, key);
            if (at >= 0) {
                double vr = ent_vec[at], vh = uvec[g];
                int n = key >> 48 ? 3 : key >> 32 ? 2 : key >> 16 ? 1 : 0;
                val[n] += (vh < vr ? vh : vr) * vr;
            }
        }
#pragma unroll
        for (int n = 0; n < CIDER_N; ++n)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val[n] += __shfl_xor_sync(0xffffffffu, val[n], o);
        if (lane == 0) {
            double delta = (double)(len_h - ref_length[r]);
            double pen = exp(-(delta * delta) / (2.0 * CIDER_SIGMA * CIDER_SIGMA));
            for (int n = 0; n < CIDER_N; ++n) {
                double nh = sqrt(norm_h[n]), nr = ref_norm[r * CIDER_N + n];
                double v = val[n];
                if (nh != 0.0 && nr != 0.0) v /= nh * nr;
                atomicAdd(&score[n], v * pen);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int n = 0; n < CIDER_N; ++n) s += score[n];
        s /= (double)CIDER_N;
        long long nref = r1 - r0;
        scores[row] = nref > 0 ? s / (double)nref * 10.0 : 0.0;
    }
}

extern "C" int ciderd_score(const void* corpus_device, const int32_t* hyp, const int32_t* video_of_row, int N, int Tc, double* scores_out,
                            unsigned long long* counts_out, s2vt_stream st) {
    if (!corpus_device || !hyp || !video_of_row || !scores_out || N <= 0 || Tc <= 0 || Tc > CIDER_MAXTOK) return S2VT_EINVAL;
    ciderd_score_kernel<<<N, 128, 0, (cudaStream_t)st>>>((const char*)corpus_device, hyp, video_of_row, Tc, scores_out, counts_out);
    return cudaGetLastError() == cudaSuccess ? S2VT_OK : S2VT_ECUDA;
}
