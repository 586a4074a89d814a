// Feature-file ingest (include/s2vt_io.h): the text format of tf_feature_extract.py:153-154 read the way
// get_video_feature_caption_pair (tf_s2vt.py:332-342) groups it, but with the file mapped once, lines indexed and numbers
// parsed by a pool of threads straight into the [n, T_v, D] float32 batch the device entry points take.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/s2vt.h"
#include "../../include/s2vt_io.h"

namespace s2vt_io {
thread_local std::string g_error;
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}
int workers(int requested) {
    int n = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : (n > 256 ? 256 : n);
}
// Run fn(worker) on n threads (the caller is worker 0).
template <class F>
void parallel(int n, F fn) {
    std::vector<std::thread> pool;
    for (int w = 1; w < n; ++w) pool.emplace_back(fn, w);
    fn(0);
    for (auto& t : pool) t.join();
}
}  // namespace s2vt_io

using s2vt_io::fail;

struct s2vt_feature_file {
    const char* text = nullptr;
    size_t len = 0;
    void* mapping = nullptr;       // non-null: we own an mmap of `len` bytes
    int32_t frames = 0, dim = 0;
    std::vector<std::string> ids;                       // video ids, order of first appearance
    std::unordered_map<std::string, int64_t> index;     // id -> video
    std::vector<uint64_t> line_of;                      // [video * frames + k] -> byte offset of that frame's line
    std::vector<uint64_t> line_end;                     // matching end offsets (exclusive, at the '\n' or EOF)
};

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

// One numeric field converted like NumPy converts a str fed as float32: float(str) (correctly rounded double; surrounding
// white space, a leading '+', inf / nan spellings allowed) then a C cast to float.
inline bool parse_field(const char* b, const char* e, float* out) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b == e) return false;
    const char* p = b;
    const bool plus = *p == '+';
    if (plus) ++p;
    if (p < e && !(plus && (*p == '-' || *p == '+'))) {       // from_chars takes '-' itself; "+-1" is not a number
        double d;
        auto r = std::from_chars(p, e, d);
        if (r.ec == std::errc() && r.ptr == e) { *out = (float)d; return true; }
        if (r.ec == std::errc::result_out_of_range && r.ptr == e) {      // float('1e999') == inf, float('1e-999') == 0.0
            std::string s(p, e);
            *out = (float)strtod(s.c_str(), nullptr);
            return true;
        }
    }
    // rare spellings Python accepts that from_chars does not ("1_000.5", "Infinity", "+nan", ...): strtod on a cleaned copy
    std::string s;
    for (const char* q = b; q < e; ++q) {
        if (*q == '_') { if (q == b || q + 1 == e || !isdigit((unsigned char)q[-1]) || !isdigit((unsigned char)q[1])) return false; continue; }
        s.push_back(*q);
    }
    if (s.find_first_of("xXpP") != std::string::npos) return false;   // C hex floats are not Python floats
    char* endp = nullptr;
    double d = strtod(s.c_str(), &endp);
    if (endp != s.c_str() + s.size() || endp == s.c_str()) return false;
    *out = (float)d;
    return true;
}

int index_text(s2vt_feature_file* f, int n_threads) {
    const char* t = f->text;
    const size_t len = f->len;
    const int nw = s2vt_io::workers(n_threads);
    // pass 1: line starts, found chunk by chunk
    std::vector<std::vector<uint64_t>> starts(nw);
    s2vt_io::parallel(nw, [&](int w) {
        size_t b = len * (size_t)w / nw, e = len * (size_t)(w + 1) / nw;
        auto& v = starts[w];
        if (w == 0 && len > 0) v.push_back(0);
        const char* p = t + b;
        while (p < t + e) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(t + e - p));
            if (!nl) break;
            if ((size_t)(nl + 1 - t) < len) v.push_back((uint64_t)(nl + 1 - t));
            p = nl + 1;
        }
    });
    std::vector<uint64_t> ls;
    for (auto& v : starts) ls.insert(ls.end(), v.begin(), v.end());
    const size_t nlines = ls.size();
    if (nlines == 0) return fail(S2VT_EINVAL, "feature file is empty");
    // pass 2 (serial, cheap: one short key per line): group by the text before the first '_' of the first field
    std::vector<std::vector<uint64_t>> per_video;      // line numbers of each video, file order
    auto line_end = [&](size_t i) { uint64_t e = i + 1 < nlines ? ls[i + 1] - 1 : len; if (e > ls[i] && t[e - 1] == '\n') --e; return e; };
    for (size_t i = 0; i < nlines; ++i) {
        const char* b = t + ls[i];
        const char* e = t + line_end(i);
        const char* comma = (const char*)memchr(b, ',', (size_t)(e - b));
        const char* fe = comma ? comma : e;
        const char* us = (const char*)memchr(b, '_', (size_t)(fe - b));
        std::string key(b, us ? us : fe);
        if (!comma && key.find_first_not_of(" \t\r") == std::string::npos)
            return fail(S2VT_EINVAL, "line %zu is blank", i + 1);
        auto it = f->index.find(key);
        int64_t v;
        if (it == f->index.end()) {
            v = (int64_t)f->ids.size();
            f->index.emplace(key, v);
            f->ids.push_back(key);
            per_video.emplace_back();
        } else v = it->second;
        per_video[(size_t)v].push_back((uint64_t)i);
    }
    f->frames = (int32_t)per_video[0].size();
    for (size_t v = 0; v < per_video.size(); ++v)
        if ((int32_t)per_video[v].size() != f->frames)      // assert len(set(feature_length)) == 1 (tf_s2vt.py:342)
            return fail(S2VT_EINVAL, "videos have different frame counts: %s has %zu, %s has %d", f->ids[v].c_str(), per_video[v].size(),
                        f->ids[0].c_str(), f->frames);
    {   // D: fields after the first of the first line
        const char* b = t + ls[0];
        const char* e = t + line_end(0);
        int commas = 0;
        for (const char* p = b; p < e; ++p) commas += *p == ',';
        f->dim = commas;
        if (commas == 0) return fail(S2VT_EINVAL, "first line has no feature fields");
    }
    f->line_of.resize(per_video.size() * (size_t)f->frames);
    f->line_end.resize(f->line_of.size());
    for (size_t v = 0; v < per_video.size(); ++v)
        for (int k = 0; k < f->frames; ++k) {
            f->line_of[v * f->frames + k] = ls[per_video[v][k]];
            f->line_end[v * f->frames + k] = line_end(per_video[v][k]);
        }
    return S2VT_OK;
}

}  // namespace

extern "C" {

const char* s2vt_io_last_error(void) { return s2vt_io::g_error.c_str(); }

int32_t s2vt_features_open_memory(const void* text, size_t len, int32_t n_threads, s2vt_feature_file** out) {
    if (!text || !out) return fail(S2VT_EINVAL, "null argument");
    s2vt_feature_file* f = new s2vt_feature_file();
    f->text = (const char*)text;
    f->len = len;
    int rc = index_text(f, n_threads);
    if (rc != S2VT_OK) { delete f; return rc; }
    *out = f;
    return S2VT_OK;
}

int32_t s2vt_features_open(const char* path, int32_t n_threads, s2vt_feature_file** out) {
    if (!path || !out) return fail(S2VT_EINVAL, "null argument");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(S2VT_ENOTFOUND, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return fail(S2VT_EINVAL, "%s is empty or unreadable", path); }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(S2VT_EINVAL, "mmap of %s failed: %s", path, strerror(errno));
    madvise(m, (size_t)st.st_size, MADV_WILLNEED);
    s2vt_feature_file* f = new s2vt_feature_file();
    f->text = (const char*)m;
    f->len = (size_t)st.st_size;
    f->mapping = m;
    int rc = index_text(f, n_threads);
    if (rc != S2VT_OK) { munmap(m, f->len); delete f; return rc; }
    *out = f;
    return S2VT_OK;
}

void s2vt_features_close(s2vt_feature_file* f) {
    if (!f) return;
    if (f->mapping) munmap(f->mapping, f->len);
    delete f;
}

int64_t s2vt_features_num_videos(const s2vt_feature_file* f) { return f ? (int64_t)f->ids.size() : 0; }
int32_t s2vt_features_num_frames(const s2vt_feature_file* f) { return f ? f->frames : 0; }
int32_t s2vt_features_dim(const s2vt_feature_file* f) { return f ? f->dim : 0; }
const char* s2vt_features_video_id(const s2vt_feature_file* f, int64_t i) {
    return (f && i >= 0 && i < (int64_t)f->ids.size()) ? f->ids[(size_t)i].c_str() : nullptr;
}
int64_t s2vt_features_find(const s2vt_feature_file* f, const char* video_id) {
    if (!f || !video_id) return -1;
    auto it = f->index.find(video_id);
    return it == f->index.end() ? -1 : it->second;
}

int32_t s2vt_features_read(const s2vt_feature_file* f, const int64_t* video_index, int64_t n, float* out, int32_t n_threads) {
    if (!f || !out || (n > 0 && !video_index) || n < 0) return fail(S2VT_EINVAL, "null argument");
    const int64_t nv = (int64_t)f->ids.size();
    for (int64_t i = 0; i < n; ++i)
        if (video_index[i] < 0 || video_index[i] >= nv) return fail(S2VT_ENOTFOUND, "video index %lld out of range [0, %lld)", (long long)video_index[i], (long long)nv);
    const int64_t items = n * f->frames;                  // one work item = one frame line
    int nw = s2vt_io::workers(n_threads);
    if (nw > items) nw = items > 0 ? (int)items : 1;
    std::atomic<int64_t> next(0);
    std::atomic<int> status(S2VT_OK);
    std::vector<std::string> errs((size_t)nw);
    const int D = f->dim, T = f->frames;
    s2vt_io::parallel(nw, [&](int w) {
        for (;;) {
            const int64_t it0 = next.fetch_add(8);
            if (it0 >= items || status.load(std::memory_order_relaxed) != S2VT_OK) return;
            const int64_t it1 = it0 + 8 < items ? it0 + 8 : items;
            for (int64_t it = it0; it < it1; ++it) {
                const int64_t slot = it / T, k = it % T;
                const size_t li = (size_t)(video_index[slot] * T + k);
                const char* p = f->text + f->line_of[li];
                const char* e = f->text + f->line_end[li];
                float* dst = out + it * D;
                const char* c = (const char*)memchr(p, ',', (size_t)(e - p));
                int j = 0;
                bool ok = c != nullptr;
                while (ok && j < D) {
                    const char* b = c + 1;
                    const char* nx = (const char*)memchr(b, ',', (size_t)(e - b));
                    const char* fe = nx ? nx : e;
                    if (!parse_field(b, fe, dst + j)) {
                        char buf[200];
                        snprintf(buf, sizeof buf, "video %s frame %lld field %d is not a number: '%.*s'", f->ids[(size_t)video_index[slot]].c_str(), (long long)k, j + 1,
                                 (int)(fe - b > 40 ? 40 : fe - b), b);
                        errs[(size_t)w] = buf;
                        status.store(S2VT_EINVAL);
                        ok = false;
                        break;
                    }
                    ++j;
                    if (!nx) { c = nullptr; break; }
                    c = nx;
                }
                if (status.load(std::memory_order_relaxed) != S2VT_OK) return;
                if (j != D || c != nullptr) {
                    char buf[200];
                    snprintf(buf, sizeof buf, "video %s frame %lld has %s than %d feature fields", f->ids[(size_t)video_index[slot]].c_str(), (long long)k,
                             j != D ? "fewer" : "more", D);
                    errs[(size_t)w] = buf;
                    status.store(S2VT_ESHAPE);
                    return;
                }
            }
        }
    });
    if (status.load() != S2VT_OK) {
        for (auto& s : errs) if (!s.empty()) return fail(status.load(), "%s", s.c_str());
        return fail(status.load(), "feature parse failed");
    }
    return S2VT_OK;
}

}  // extern "C"
