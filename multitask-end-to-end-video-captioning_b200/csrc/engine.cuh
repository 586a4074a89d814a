// Handle layout and helpers shared by the translation units of libs2vt_b200.so.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/s2vt.h"
#include "common.cuh"

struct Var {
    std::string name;
    std::vector<std::string> aliases;
    int64_t rows, cols;  // cols == 0 -> 1-D
    size_t off;          // floats
    size_t count() const { return (size_t)rows * (cols ? cols : 1); }
};

// Bump allocator over a caller-owned block; with base == nullptr it only measures.
struct Arena {
    char* base;
    size_t cap, used;
    bool overflow;
    Arena(void* b, size_t c) : base((char*)b), cap(c), used(0), overflow(false) {}
    template <typename U> U* take(size_t n) {
        size_t bytes = ru64(n * sizeof(U), 256);
        size_t o = used;
        used += bytes;
        if (base && used > cap) { overflow = true; return nullptr; }
        return base ? reinterpret_cast<U*>(base + o) : nullptr;
    }
};

struct s2vt_handle {
    s2vt_config cfg;
    int D, E, H, V, Tv, Tc, A, T;        // logical dims, T = Tv + Tc
    int Dp, Ep, Hp, Vp, Gp, Ap;          // padded dims, Gp = 4*Hp
    size_t esz;                           // sizeof(compute dtype)
    std::vector<Var> vars;
    size_t P;                             // floats in the flat parameter vector
    // state block
    char* state; size_t state_bytes;
    char* ws; size_t ws_bytes;
    float *params, *grads, *adam_m, *adam_v;
    double* sq;                           // [0] dense grad sumsq, [1] dense Wemb sumsq, [2] Wemb slice sumsq, [3] scratch
    float* scal;                          // small fp32 scratch (loss parts, norm)
    unsigned* gbar;                       // grid-barrier counters of the persistent step chains ([0] caller's stream, [16] side stream)
    void *WeT, *W1xT, *W1hT, *W1h, *W1x, *W2xT, *W2x, *W2eT, *W2e, *W2hT, *W2h, *WoT, *Wo, *WembC, *attrWT;
    float *be_p, *b1_p, *b2_p, *bo_p, *Etab;
    bool bound, fresh;
    // instrumentation (bench.py): launch counter and optional CUDA-event brackets around GEMM launches
    mutable long long launches = 0;
    bool prof = false;
    struct ProfRec { cudaEvent_t a, b; double flops, bytes; int cls; int M, N, K; long long count, launches; };
    // an open chain of per-step launches (bracketed as a whole so programmatic dependent launches keep overlapping)
    bool chain_open = false; ProfRec chain;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    void* tc_cache = nullptr;
    // LSTM1 forward cache shared by s2vt_rollout and the following training call (opt-in, s2vt_set_reuse_frontend)
    bool reuse_front = false, front_valid = false;
    int front_B = 0; const float* front_video = nullptr;
    // internal side stream for the LSTM1 backward chain (fork/join inside one call; invisible to the caller)
    cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_refresh = nullptr, ev_wo = nullptr;
    bool wo_grad_early = false;           // the last backward recorded ev_wo when d embed_word_W / d embed_word_b became final (s2vt_grad_segment_ready)
    // data-parallel exchange over NVLink peer memory (peer.cuh): mapped gradient blocks / flag blocks of all ranks
    int peer_rank = -1, peer_world = 0, peer_grid = 0; unsigned peer_epoch = 0;
    void* peer_comm = nullptr; void* peer_opened[32] = {nullptr}; char* peer_s[16] = {nullptr}; void* peer_c[16] = {nullptr};   // state blocks / flag blocks of all ranks
    bool opt_sharded = false;             // the Adam slots are current in this rank's slice only (s2vt_peer_optimizer_step; s2vt_peer_gather_state clears it)
    cudaEvent_t ev_gate = nullptr;        // the gated side-stream part of dout1 (rows the LSTM2 BPTT chain finished first) is complete
    cudaEvent_t ev_seg[2] = {nullptr, nullptr};   // ... and these when d Wemb (segment 1) / d LSTM2 weights + biases (segment 2) became final
    unsigned seg_ready = 0;               // bit i: segment i of the last backward may be handed out
    bool copies_zeroed = false;
    int overlap = 135;                    // bit 0: late refresh, bit 1: dWo, bit 2: LSTM1 backward chain run on the side stream; bit 7: gated consumption of dG2 under the LSTM2 BPTT chain           // tc::MapCache (TMA tensor maps keyed by pointer / shape)
    // variable indices
    int iWemb, iWe, ibe, iWo, ibo, iW1, ib1, iW2, ib2, iAW, iAb;
    mutable std::string err;
    int fail(int code, const char* fmt, ...) const {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
    float* P_(int i) const { return params + vars[i].off; }
    float* G_(int i) const { return grads + vars[i].off; }
};

#define CUDA_TRY(h, expr)                                                                                        \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess) return (h)->fail(S2VT_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define KCHECK(h) do { (h)->launches++; CUDA_TRY(h, cudaGetLastError()); } while (0)
