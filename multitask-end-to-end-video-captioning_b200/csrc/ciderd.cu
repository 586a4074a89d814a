// placeholder -- replaced by the GPU CIDEr-D scorer
#include "engine.cuh"
extern "C" int ciderd_corpus_create(const int32_t*, const int64_t*, int64_t, const int64_t*, int64_t, ciderd_corpus**) { return S2VT_ESTATE; }
extern "C" void ciderd_corpus_destroy(ciderd_corpus*) {}
extern "C" size_t ciderd_corpus_device_bytes(const ciderd_corpus*) { return 0; }
extern "C" int ciderd_corpus_serialize(const ciderd_corpus*, void*) { return S2VT_ESTATE; }
extern "C" int ciderd_score(const void*, const int32_t*, const int32_t*, int, int, double*, unsigned long long*, s2vt_stream) { return S2VT_ESTATE; }
