// TensorFlow checkpoint reader without TensorFlow (include/s2vt_io.h).
//
// The reference saves with tf.train.Saver(max_to_keep=100, write_version=1) (final_beam_search.py:447, e2e_beam_search.py:525,
// ...) or the default V2 writer (reinforcement_multisampling_tf_s2vt.py) and restores through optimistic_restore
// (:47-61), which needs {variable name -> shape} and the tensors.  Both formats sit on TensorFlow's SSTable
// (tensorflow/core/lib/io/table*, the LevelDB table format):
//   file   = data blocks | metaindex block | index block | footer(48 B: two BlockHandles, padding, magic 0xdb4775248b80fb57)
//   block  = entries | restart offsets (u32 each) | num_restarts (u32) ; followed by 1 type byte (0 = raw) + masked CRC32C
//   entry  = varint shared, varint non_shared, varint value_len, key suffix, value (keys are prefix-compressed)
// V2 "<prefix>.index": key "" -> BundleHeaderProto, key <name> -> BundleEntryProto{dtype, shape, shard_id, offset, size,
//    crc32c}; tensor bytes live in "<prefix>.data-SSSSS-of-NNNNN" (tensorflow/core/util/tensor_bundle).
// V1 "<prefix>": key "" -> SavedTensorSlices{meta}, other keys -> SavedTensorSlices{data = SavedSlice{name, slice,
//    TensorProto}} with the values in the typed repeated field (float_val ...) (tensorflow/core/util/saved_tensor_slice.proto).
// PARITY UNPINNED: no TensorFlow and no checkpoint file exists in this environment; the tests write both formats with an
// independent Python writer of the same published layout (tests/tf_ckpt_writer.py).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/s2vt.h"
#include "../../include/s2vt_io.h"

namespace s2vt_io {
int fail(int code, const char* fmt, ...);
}
using s2vt_io::fail;

namespace {

// ---- CRC32C (Castagnoli), slicing-by-8 -----------------------------------------------------------------------------
struct Crc32cTable {
    uint32_t t[8][256];
    Crc32cTable() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    }
};
uint32_t crc32c(const uint8_t* p, size_t n, uint32_t crc = 0) {
    static const Crc32cTable T;
    crc = ~crc;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= crc;
        crc = T.t[7][w & 0xFF] ^ T.t[6][(w >> 8) & 0xFF] ^ T.t[5][(w >> 16) & 0xFF] ^ T.t[4][(w >> 24) & 0xFF] ^ T.t[3][(w >> 32) & 0xFF] ^
              T.t[2][(w >> 40) & 0xFF] ^ T.t[1][(w >> 48) & 0xFF] ^ T.t[0][w >> 56];
        p += 8; n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xFF];
    return ~crc;
}
inline uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }   // crc32c::Mask

// ---- little helpers ----------------------------------------------------------------------------------------------
struct Span { const uint8_t* p = nullptr; size_t n = 0; };

bool get_varint(const uint8_t*& p, const uint8_t* e, uint64_t* v) {
    uint64_t r = 0;
    for (int shift = 0; shift < 64 && p < e; shift += 7) {
        uint8_t b = *p++;
        r |= (uint64_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) { *v = r; return true; }
    }
    return false;
}
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

// One protobuf field: number, wire type, varint / fixed value or length-delimited span.
struct Field { uint32_t num = 0, wt = 0; uint64_t val = 0; Span bytes; };
bool next_field(const uint8_t*& p, const uint8_t* e, Field* f) {
    uint64_t tag;
    if (!get_varint(p, e, &tag)) return false;
    f->num = (uint32_t)(tag >> 3); f->wt = (uint32_t)(tag & 7); f->val = 0; f->bytes = Span();
    switch (f->wt) {
        case 0: return get_varint(p, e, &f->val);
        case 1: if (e - p < 8) return false; memcpy(&f->val, p, 8); f->bytes = {p, 8}; p += 8; return true;
        case 5: if (e - p < 4) return false; f->val = rd32(p); f->bytes = {p, 4}; p += 4; return true;
        case 2: { uint64_t n; if (!get_varint(p, e, &n) || (uint64_t)(e - p) < n) return false; f->bytes = {p, (size_t)n}; p += n; return true; }
        default: return false;
    }
}

struct Mapped {
    const uint8_t* p = nullptr; size_t n = 0;
    int open(const std::string& path) {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return fail(S2VT_ENOTFOUND, "cannot open %s", path.c_str());
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size == 0) { ::close(fd); return fail(S2VT_EINVAL, "%s is empty", path.c_str()); }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) return fail(S2VT_EINVAL, "mmap of %s failed", path.c_str());
        p = (const uint8_t*)m; n = (size_t)st.st_size;
        return S2VT_OK;
    }
    void close() { if (p) munmap((void*)p, n); p = nullptr; n = 0; }
};

// ---- SSTable -------------------------------------------------------------------------------------------------------
const uint64_t kTableMagic = 0xdb4775248b80fb57ull;

int read_block(const Mapped& f, uint64_t off, uint64_t size, Span* out) {
    if (f.n < 5 || size > f.n - 5 || off > f.n - 5 - size) return fail(S2VT_EINVAL, "table block [%llu,+%llu) runs past the end of the file", (unsigned long long)off, (unsigned long long)size);
    const uint8_t* b = f.p + off;
    if (b[size] != 0) return fail(S2VT_EINVAL, "table block is compressed (type %d); TF Savers write raw blocks", (int)b[size]);
    if (crc_mask(crc32c(b, size + 1)) != rd32(b + size + 1)) return fail(S2VT_EINVAL, "table block at %llu fails its CRC32C", (unsigned long long)off);
    *out = {b, (size_t)size};
    return S2VT_OK;
}

// Calls fn(key, value) for every entry of a block; fn returns an S2VT code.
template <class F>
int for_each_entry(Span blk, F fn) {
    if (blk.n < 4) return fail(S2VT_EINVAL, "table block too small");
    const uint32_t nrestart = rd32(blk.p + blk.n - 4);
    if ((uint64_t)nrestart * 4 + 4 > blk.n) return fail(S2VT_EINVAL, "bad restart array");
    const uint8_t* p = blk.p;
    const uint8_t* e = blk.p + blk.n - 4 - (size_t)nrestart * 4;
    std::string key;
    while (p < e) {
        uint64_t shared, non_shared, vlen;
        if (!get_varint(p, e, &shared) || !get_varint(p, e, &non_shared) || !get_varint(p, e, &vlen) || shared > key.size() ||
            (uint64_t)(e - p) < non_shared + vlen)
            return fail(S2VT_EINVAL, "corrupt table entry");
        key.resize((size_t)shared);
        key.append((const char*)p, (size_t)non_shared);
        p += non_shared;
        int rc = fn(key, Span{p, (size_t)vlen});
        if (rc != S2VT_OK) return rc;
        p += vlen;
    }
    return S2VT_OK;
}

template <class F>
int for_each_table_entry(const Mapped& f, F fn) {
    if (f.n < 48) return fail(S2VT_EINVAL, "file too small for an SSTable footer");
    const uint8_t* foot = f.p + f.n - 48;
    uint64_t magic;
    memcpy(&magic, foot + 40, 8);
    if (magic != kTableMagic) return fail(S2VT_EINVAL, "not a TensorFlow table file (bad magic)");
    const uint8_t* p = foot;
    uint64_t mo, ms, io, is;
    if (!get_varint(p, foot + 40, &mo) || !get_varint(p, foot + 40, &ms) || !get_varint(p, foot + 40, &io) || !get_varint(p, foot + 40, &is))
        return fail(S2VT_EINVAL, "corrupt table footer");
    Span index;
    int rc = read_block(f, io, is, &index);
    if (rc != S2VT_OK) return rc;
    return for_each_entry(index, [&](const std::string&, Span handle) {
        const uint8_t* q = handle.p;
        uint64_t bo, bs;
        if (!get_varint(q, handle.p + handle.n, &bo) || !get_varint(q, handle.p + handle.n, &bs)) return fail(S2VT_EINVAL, "corrupt index entry");
        Span data;
        int r = read_block(f, bo, bs, &data);
        if (r != S2VT_OK) return r;
        return for_each_entry(data, fn);
    });
}

// ---- protos ------------------------------------------------------------------------------------------------------
bool parse_shape(Span s, std::vector<int64_t>* dims) {          // TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
    const uint8_t* p = s.p; const uint8_t* e = s.p + s.n;
    Field f;
    while (p < e) {
        if (!next_field(p, e, &f)) return false;
        if (f.num == 2 && f.wt == 2) {
            const uint8_t* q = f.bytes.p; const uint8_t* qe = q + f.bytes.n;
            Field d; int64_t size = 0;
            while (q < qe) { if (!next_field(q, qe, &d)) return false; if (d.num == 1 && d.wt == 0) size = (int64_t)d.val; }
            dims->push_back(size);
        } else if (f.num == 3 && f.wt == 0 && f.val) return false;   // unknown_rank
    }
    return true;
}

// TensorSliceProto: repeated Extent extent = 1 {int64 start = 1; oneof {int64 length = 2}}.  Full iff every extent has no
// length (the "-" spec) or covers [0, dim).
bool slice_is_full(Span s, const std::vector<int64_t>& dims) {
    const uint8_t* p = s.p; const uint8_t* e = s.p + s.n;
    Field f; size_t i = 0;
    while (p < e) {
        if (!next_field(p, e, &f)) return false;
        if (f.num != 1 || f.wt != 2) continue;
        const uint8_t* q = f.bytes.p; const uint8_t* qe = q + f.bytes.n;
        Field x; int64_t start = 0, length = -1;
        while (q < qe) { if (!next_field(q, qe, &x)) return false; if (x.num == 1) start = (int64_t)x.val; else if (x.num == 2) length = (int64_t)x.val; }
        if (i >= dims.size()) return false;
        if (!(start == 0 && (length < 0 || length == dims[i]))) return false;
        ++i;
    }
    return i == dims.size();
}

struct Tensor {
    std::string name;
    int32_t dtype = 0;
    std::vector<int64_t> dims;
    // V2
    int32_t shard = 0; int64_t offset = 0, size = 0; uint32_t crc = 0;
    // V1
    Span proto;            // TensorProto bytes inside the mapped file
    bool full = true;      // single full slice
    int64_t count() const { int64_t n = 1; for (int64_t d : dims) n *= d; return n; }
};

size_t dtype_size(int32_t dt) { return dt == S2VT_DT_FLOAT || dt == S2VT_DT_INT32 ? 4 : (dt == S2VT_DT_DOUBLE || dt == S2VT_DT_INT64 ? 8 : 0); }

template <class T>
void convert(const uint8_t* src, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) { T v; memcpy(&v, src + i * sizeof(T), sizeof(T)); out[i] = (float)v; }
}

}  // namespace

struct s2vt_ckpt {
    int format = 0;
    Mapped table;
    std::vector<Mapped> shards;
    std::vector<Tensor> tensors;           // key order
    std::map<std::string, int> by_name;
    ~s2vt_ckpt() { table.close(); for (auto& s : shards) s.close(); }
};

namespace {

int open_v2(s2vt_ckpt* c, const std::string& prefix) {
    int num_shards = 1;
    int rc = for_each_table_entry(c->table, [&](const std::string& key, Span val) {
        const uint8_t* p = val.p; const uint8_t* e = val.p + val.n;
        Field f;
        if (key.empty()) {                  // BundleHeaderProto {int32 num_shards = 1; Endianness endianness = 2; VersionDef version = 3}
            while (p < e) {
                if (!next_field(p, e, &f)) return fail(S2VT_EINVAL, "corrupt bundle header");
                if (f.num == 1) num_shards = (int)f.val;
                if (f.num == 2 && f.val != 0) return fail(S2VT_EINVAL, "big-endian bundle");
            }
            return (int)S2VT_OK;
        }
        Tensor t; t.name = key;
        while (p < e) {                     // BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, fixed32 crc32c=6, slices=7}
            if (!next_field(p, e, &f)) return fail(S2VT_EINVAL, "corrupt bundle entry for %s", key.c_str());
            switch (f.num) {
                case 1: t.dtype = (int32_t)f.val; break;
                case 2: if (!parse_shape(f.bytes, &t.dims)) return fail(S2VT_EINVAL, "bad shape for %s", key.c_str()); break;
                case 3: t.shard = (int32_t)f.val; break;
                case 4: t.offset = (int64_t)f.val; break;
                case 5: t.size = (int64_t)f.val; break;
                case 6: t.crc = (uint32_t)f.val; break;
                case 7: t.full = false; break;      // partitioned variable: slices are stored under other keys
                default: break;
            }
        }
        c->tensors.push_back(t);
        return (int)S2VT_OK;
    });
    if (rc != S2VT_OK) return rc;
    if (num_shards < 1 || num_shards > 99999) return fail(S2VT_EINVAL, "bad shard count %d", num_shards);
    c->shards.resize((size_t)num_shards);
    for (int s = 0; s < num_shards; ++s) {
        char suffix[64];
        snprintf(suffix, sizeof suffix, ".data-%05d-of-%05d", s, num_shards);
        rc = c->shards[(size_t)s].open(prefix + suffix);
        if (rc != S2VT_OK) return rc;
    }
    return S2VT_OK;
}

int open_v1(s2vt_ckpt* c) {
    std::map<std::string, Tensor> meta;
    std::map<std::string, std::pair<int, Span>> data;      // name -> (#slices seen, TensorProto of the last)
    int rc = for_each_table_entry(c->table, [&](const std::string& key, Span val) {
        const uint8_t* p = val.p; const uint8_t* e = val.p + val.n;
        Field f;
        while (p < e) {                     // SavedTensorSlices {SavedTensorSliceMeta meta = 1; SavedSlice data = 2}
            if (!next_field(p, e, &f)) return fail(S2VT_EINVAL, "corrupt SavedTensorSlices");
            if (f.num == 1 && f.wt == 2 && key.empty()) {
                const uint8_t* q = f.bytes.p; const uint8_t* qe = q + f.bytes.n;
                Field m;
                while (q < qe) {            // SavedTensorSliceMeta {repeated SavedSliceMeta tensor = 1; VersionDef versions = 2}
                    if (!next_field(q, qe, &m)) return fail(S2VT_EINVAL, "corrupt SavedTensorSliceMeta");
                    if (m.num != 1 || m.wt != 2) continue;
                    const uint8_t* r = m.bytes.p; const uint8_t* re = r + m.bytes.n;
                    Field x; Tensor t; int nslices = 0; std::vector<Span> slices;
                    while (r < re) {        // SavedSliceMeta {name = 1; TensorShapeProto shape = 2; DataType type = 3; repeated TensorSliceProto slice = 4}
                        if (!next_field(r, re, &x)) return fail(S2VT_EINVAL, "corrupt SavedSliceMeta");
                        if (x.num == 1) t.name.assign((const char*)x.bytes.p, x.bytes.n);
                        else if (x.num == 2) { if (!parse_shape(x.bytes, &t.dims)) return fail(S2VT_EINVAL, "bad shape in meta"); }
                        else if (x.num == 3) t.dtype = (int32_t)x.val;
                        else if (x.num == 4) { ++nslices; slices.push_back(x.bytes); }
                    }
                    t.full = nslices == 1 && slice_is_full(slices[0], t.dims);
                    meta[t.name] = t;
                }
            } else if (f.num == 2 && f.wt == 2) {
                const uint8_t* q = f.bytes.p; const uint8_t* qe = q + f.bytes.n;
                Field x; std::string name; Span proto;
                while (q < qe) {            // SavedSlice {name = 1; TensorSliceProto slice = 2; TensorProto data = 3}
                    if (!next_field(q, qe, &x)) return fail(S2VT_EINVAL, "corrupt SavedSlice");
                    if (x.num == 1) name.assign((const char*)x.bytes.p, x.bytes.n);
                    else if (x.num == 3) proto = x.bytes;
                }
                auto& d = data[name];
                d.first += 1; d.second = proto;
            }
        }
        return (int)S2VT_OK;
    });
    if (rc != S2VT_OK) return rc;
    if (meta.empty()) return fail(S2VT_EINVAL, "V1 checkpoint has no SavedTensorSliceMeta entry");
    for (auto& kv : meta) {
        Tensor t = kv.second;
        auto it = data.find(t.name);
        if (it == data.end()) return fail(S2VT_EINVAL, "V1 checkpoint lists %s but stores no slice for it", t.name.c_str());
        if (it->second.first != 1) t.full = false;
        t.proto = it->second.second;
        c->tensors.push_back(t);
    }
    return S2VT_OK;
}

// Values of a V1 TensorProto as float32.  The V1 writer fills the typed repeated field (packed); tensor_content is accepted too.
int read_v1(const Tensor& t, float* out) {
    const int64_t n = t.count();
    const uint8_t* p = t.proto.p; const uint8_t* e = p + t.proto.n;
    const uint32_t want = t.dtype == S2VT_DT_FLOAT ? 5 : t.dtype == S2VT_DT_DOUBLE ? 6 : t.dtype == S2VT_DT_INT32 ? 7 : 10;
    int64_t got = 0;
    Field f;
    while (p < e) {
        if (!next_field(p, e, &f)) return fail(S2VT_EINVAL, "corrupt TensorProto for %s", t.name.c_str());
        if (f.num == 4 && f.wt == 2 && f.bytes.n) {                       // tensor_content
            if ((int64_t)f.bytes.n != n * (int64_t)dtype_size(t.dtype)) return fail(S2VT_ESHAPE, "%s: tensor_content has %zu bytes", t.name.c_str(), f.bytes.n);
            if (t.dtype == S2VT_DT_FLOAT) memcpy(out, f.bytes.p, f.bytes.n);
            else if (t.dtype == S2VT_DT_DOUBLE) convert<double>(f.bytes.p, n, out);
            else if (t.dtype == S2VT_DT_INT32) convert<int32_t>(f.bytes.p, n, out);
            else convert<int64_t>(f.bytes.p, n, out);
            got = n;
        } else if (f.num == want) {
            if (f.wt == 2) {                                                // packed
                if (want == 5 || want == 6) {
                    const size_t es = want == 5 ? 4 : 8;
                    const int64_t k = (int64_t)(f.bytes.n / es);
                    if (got + k > n) return fail(S2VT_ESHAPE, "%s holds more values than its shape", t.name.c_str());
                    if (want == 5) memcpy(out + got, f.bytes.p, (size_t)k * 4); else convert<double>(f.bytes.p, k, out + got);
                    got += k;
                } else {
                    const uint8_t* q = f.bytes.p; const uint8_t* qe = q + f.bytes.n;
                    while (q < qe) {
                        uint64_t v;
                        if (!get_varint(q, qe, &v) || got >= n) return fail(S2VT_ESHAPE, "%s: bad packed integers", t.name.c_str());
                        out[got++] = want == 7 ? (float)(int32_t)v : (float)(int64_t)v;
                    }
                }
            } else {                                                        // one unpacked element
                if (got >= n) return fail(S2VT_ESHAPE, "%s holds more values than its shape", t.name.c_str());
                // wire types: float = 5 (fixed32), double = 1 (fixed64), integers = 0 (varint); anything else is a malformed record
                if ((want == 5 && f.wt != 5) || (want == 6 && f.wt != 1) || (want != 5 && want != 6 && f.wt != 0))
                    return fail(S2VT_EINVAL, "%s: element with wire type %d does not match its dtype", t.name.c_str(), (int)f.wt);
                if (want == 5) { float v; memcpy(&v, f.bytes.p, 4); out[got++] = v; }
                else if (want == 6) { double v; memcpy(&v, f.bytes.p, 8); out[got++] = (float)v; }
                else out[got++] = want == 7 ? (float)(int32_t)f.val : (float)(int64_t)f.val;
            }
        }
    }
    if (got == 1 && n > 1) { for (int64_t i = 1; i < n; ++i) out[i] = out[0]; got = n; }   // TensorProto splat convention
    if (got != n) return fail(S2VT_ESHAPE, "%s: %lld values stored, shape needs %lld", t.name.c_str(), (long long)got, (long long)n);
    return S2VT_OK;
}

bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }

}  // namespace

extern "C" {

int32_t s2vt_ckpt_open(const char* prefix, s2vt_ckpt** out) {
    if (!prefix || !out) return fail(S2VT_EINVAL, "null argument");
    std::string pre(prefix);
    s2vt_ckpt* c = new s2vt_ckpt();
    int rc;
    if (exists(pre + ".index")) {
        c->format = 2;
        rc = c->table.open(pre + ".index");
        if (rc == S2VT_OK) rc = open_v2(c, pre);
    } else if (exists(pre)) {
        c->format = 1;
        rc = c->table.open(pre);
        if (rc == S2VT_OK) rc = open_v1(c);
    } else {
        rc = fail(S2VT_ENOTFOUND, "no checkpoint at %s (neither %s.index nor %s)", prefix, prefix, prefix);
    }
    if (rc != S2VT_OK) { delete c; return rc; }
    for (size_t i = 0; i < c->tensors.size(); ++i) c->by_name[c->tensors[i].name] = (int)i;
    *out = c;
    return S2VT_OK;
}

void s2vt_ckpt_close(s2vt_ckpt* c) { delete c; }
int32_t s2vt_ckpt_format(const s2vt_ckpt* c) { return c ? c->format : 0; }
int32_t s2vt_ckpt_num_tensors(const s2vt_ckpt* c) { return c ? (int32_t)c->tensors.size() : 0; }

int32_t s2vt_ckpt_tensor_info(const s2vt_ckpt* c, int32_t i, const char** name, int32_t* dtype, int32_t* ndim, int64_t* dims) {
    if (!c || i < 0 || i >= (int32_t)c->tensors.size()) return fail(S2VT_EINVAL, "tensor index out of range");
    const Tensor& t = c->tensors[(size_t)i];
    if (t.dims.size() > 8) return fail(S2VT_ESHAPE, "%s has rank %zu > 8", t.name.c_str(), t.dims.size());
    if (name) *name = t.name.c_str();
    if (dtype) *dtype = t.dtype;
    if (ndim) *ndim = (int32_t)t.dims.size();
    if (dims) for (size_t k = 0; k < t.dims.size(); ++k) dims[k] = t.dims[k];
    return S2VT_OK;
}

int32_t s2vt_ckpt_find(const s2vt_ckpt* c, const char* name) {
    if (!c || !name) return -1;
    auto it = c->by_name.find(name);
    return it == c->by_name.end() ? -1 : it->second;
}

int32_t s2vt_ckpt_read_f32(const s2vt_ckpt* c, int32_t i, float* out, int64_t capacity) {
    if (!c || !out || i < 0 || i >= (int32_t)c->tensors.size()) return fail(S2VT_EINVAL, "bad argument");
    const Tensor& t = c->tensors[(size_t)i];
    const int64_t n = t.count();
    if (capacity < n) return fail(S2VT_ENOSPACE, "%s needs %lld floats, buffer holds %lld", t.name.c_str(), (long long)n, (long long)capacity);
    if (!dtype_size(t.dtype)) return fail(S2VT_EINVAL, "%s has dtype %d (only float / double / int32 / int64 are read)", t.name.c_str(), t.dtype);
    if (!t.full) return fail(S2VT_EINVAL, "%s is stored as partitioned slices", t.name.c_str());
    if (c->format == 1) return read_v1(t, out);
    if (t.shard < 0 || t.shard >= (int32_t)c->shards.size()) return fail(S2VT_EINVAL, "%s names shard %d", t.name.c_str(), t.shard);
    const Mapped& m = c->shards[(size_t)t.shard];
    if (t.offset < 0 || t.size != n * (int64_t)dtype_size(t.dtype) || (uint64_t)t.offset + (uint64_t)t.size > m.n)
        return fail(S2VT_ESHAPE, "%s: entry [%lld,+%lld) does not fit its shape / data file", t.name.c_str(), (long long)t.offset, (long long)t.size);
    const uint8_t* src = m.p + t.offset;
    if (crc_mask(crc32c(src, (size_t)t.size)) != t.crc) return fail(S2VT_EINVAL, "%s fails its CRC32C", t.name.c_str());
    if (t.dtype == S2VT_DT_FLOAT) memcpy(out, src, (size_t)t.size);
    else if (t.dtype == S2VT_DT_DOUBLE) convert<double>(src, n, out);
    else if (t.dtype == S2VT_DT_INT32) convert<int32_t>(src, n, out);
    else convert<int64_t>(src, n, out);
    return S2VT_OK;
}

}  // extern "C"
