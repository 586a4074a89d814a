// placeholder -- replaced by the batched device beam search
#include "engine.cuh"
extern "C" int s2vt_beam_search(s2vt_handle* h, const float*, int, int, float, int32_t*, int32_t*, float*, float*, s2vt_stream) { return h->fail(S2VT_ESTATE, "beam search not built yet"); }
extern "C" int s2vt_beam_init(s2vt_handle* h, const float*, float*, float*, s2vt_stream) { return h->fail(S2VT_ESTATE, "beam search not built yet"); }
extern "C" int s2vt_beam_step(s2vt_handle* h, const float*, const float*, const int32_t*, int, int32_t*, float*, float*, float*, s2vt_stream) { return h->fail(S2VT_ESTATE, "beam search not built yet"); }
