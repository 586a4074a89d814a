// Shared device helpers for the S2VT B200 library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;
typedef __half f16;

// Operand type of the FORWARD GEMMs of a compute mode.  The bf16 mode (tensor-core mode) keeps every forward value -- activations,
// the K-major weight copies, the word-embedding table -- in IEEE fp16: same 16 bits and the same tcgen05 kind::f16 rate as bf16,
// but 10 mantissa bits instead of 7, which is what brings the teacher-forced logits within the 1e-3 the north star quotes
// (bf16 operands: 4.4e-3 at H = 1000, tests/test_gpu_bench_config_parity.py).  Every forward quantity is bounded (|h| < 1,
// post-ReLU features, weights of magnitude 0.1), far inside fp16's range.  Gradients keep bf16: REINFORCE gate gradients reach
// 1e-9 and below, under fp16's normal range.  The weight-gradient GEMMs X^T . dY round their activation operand to bf16 first
// (kind::f16 takes one format for both operands; the mixed descriptor faults on B200).
template <typename T> struct Fwd { typedef T type; };
template <> struct Fwd<bf16> { typedef f16 type; };

#define S2VT_PAD 128  // every contraction / output dimension is zero-padded to a multiple of this

__host__ __device__ inline int ru(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t ru64(size_t x, size_t m) { return (x + m - 1) / m * m; }

template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ f16 from_f32<f16>(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f32(f16 x) { return __half2float(x); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10.  Must match oracle/philox.py bit for bit (Random123 KAT in tests/test_oracle_model.py).
//   sampler : c0 = vocab index / 4, c1 = decode step, c2 = global row, c3 = STREAM_SAMPLE
//   dropout : c0 = unit / 4,        c1 = time step,   c2 = global row, c3 = STREAM_DROP1/2
// ---------------------------------------------------------------------------------------------
#define S2VT_STREAM_SAMPLE 0x53414D50u
#define S2VT_STREAM_DROP1 0x44525031u
#define S2VT_STREAM_DROP2 0x44525032u

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u32_to_uniform(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f; }

// Dropout multiplier (0 or 1/keep) for unit u of global row `row` at time step `step`.
__device__ __forceinline__ float dropout_mult(unsigned long long seed, uint32_t stream, uint32_t row, uint32_t step, uint32_t u, float keep) {
    uint4 o = philox4x32_10(u >> 2, step, row, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t x = (u & 3) == 0 ? o.x : (u & 3) == 1 ? o.y : (u & 3) == 2 ? o.z : o.w;
    return u32_to_uniform(x) < keep ? 1.0f / keep : 0.0f;
}
// Four consecutive units (u % 4 == 0) at once.
__device__ __forceinline__ float4 dropout_mult4(unsigned long long seed, uint32_t stream, uint32_t row, uint32_t step, uint32_t u, float keep) {
    uint4 o = philox4x32_10(u >> 2, step, row, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
    float ik = 1.0f / keep;
    return make_float4(u32_to_uniform(o.x) < keep ? ik : 0.f, u32_to_uniform(o.y) < keep ? ik : 0.f,
                       u32_to_uniform(o.z) < keep ? ik : 0.f, u32_to_uniform(o.w) < keep ? ik : 0.f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// Gate non-linearities by compute dtype: fp32 mode uses the accurate library functions; bf16 mode the MUFU paths
// (ex2 + rcp, absolute error ~1e-6, far below the bf16 rounding of the operands).
template <typename T> __device__ __forceinline__ float sigm(float x);
template <> __device__ __forceinline__ float sigm<float>(float x) { return 1.0f / (1.0f + expf(-x)); }
template <> __device__ __forceinline__ float sigm<bf16>(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
template <> __device__ __forceinline__ float sigm<f16>(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
template <typename T> __device__ __forceinline__ float tanh_(float x);
template <> __device__ __forceinline__ float tanh_<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ float tanh_<bf16>(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
template <> __device__ __forceinline__ float tanh_<f16>(float x) { return 2.0f * __fdividef(1.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
// four saved gate activations (si, tj, sf, so) in the compute dtype
__device__ __forceinline__ float4 load_gates4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load_gates4(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 load_gates4(const f16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
// the same four activations kept packed (as stored) while they wait in registers: 2 registers instead of 4 in bf16 mode
template <typename T> struct GateRaw;
template <> struct GateRaw<float> { float4 v; };
template <> struct GateRaw<bf16> { uint2 v; };
template <> struct GateRaw<f16> { uint2 v; };
__device__ __forceinline__ void load_gates_raw(const float* p, GateRaw<float>& r) { r.v = *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void load_gates_raw(const bf16* p, GateRaw<bf16>& r) { r.v = *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void load_gates_raw(const f16* p, GateRaw<f16>& r) { r.v = *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ float4 unpack_gates(const GateRaw<f16>& r) {
    float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 unpack_gates(const GateRaw<float>& r) { return r.v; }
__device__ __forceinline__ float4 unpack_gates(const GateRaw<bf16>& r) {
    float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.v.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store_gates4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store_gates4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ void store_gates4(f16* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// Gumbel(0,1) noise -log(-log(u)) on the MUFU log path.  E = -log(u) is the delicate part for u -> 1 (the large noise
// values that win the arg-max): there 1-u is exact in fp32 and a 3-term series of -log1p(-t) is used (rel. err < 1e-5).
__device__ __forceinline__ float gumbel_fast(float u) {
    float t = 1.0f - u;
    float E = t < 0.03125f ? t * (1.0f + t * (0.5f + t * 0.33333334f)) : -__logf(u);
    return -__logf(E);
}

// Block-wide reductions (blockDim.x multiple of 32, <= 1024).  `red` is a shared array of >= 32 elements.
template <typename V, class Op>
__device__ __forceinline__ V block_reduce(V v, Op op, V* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    V r = red[0];
    for (int i = 1; i < nw; ++i) r = op(r, red[i]);
    return r;
}

struct ArgVal {
    float v;
    int i;
};
// Larger value wins; ties go to the LOWER index (tf.argmax / tf.nn.top_k convention).
__device__ __forceinline__ ArgVal argmax_op(ArgVal a, ArgVal b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }
__device__ __forceinline__ ArgVal block_argmax(ArgVal v, ArgVal* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgVal t;
        t.v = __shfl_xor_sync(0xffffffffu, v.v, o);
        t.i = __shfl_xor_sync(0xffffffffu, v.i, o);
        v = argmax_op(v, t);
    }
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    ArgVal r = red[0];
    for (int i = 1; i < nw; ++i) r = argmax_op(r, red[i]);
    return r;
}
