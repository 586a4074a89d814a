// Microbenchmark (debug aid): cycles per tcgen05.mma (M=128, K=16, bf16, operands in shared memory) as a function of N and of the
// number of independent accumulator tiles, for a back-to-back K loop issued by one thread (what a recurrent-step kernel does).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }

template <int N>
__global__ void __launch_bounds__(64) rate(int nmma, int nacc, int kblocks_distinct, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + N * 128) * (kblocks_distinct < 0 ? 1 : kblocks_distinct) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x >= 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    const uint32_t tmem = slot;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 32) {
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            if (kblocks_distinct < 0) {      // lean issue loop: no divisions, operands fixed, accumulate always on after the first
                const uint32_t a = s32(smem);
                const uint64_t ad = make_desc(a), bd = make_desc(a + 16384);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(IDESC), "r"(0) : "memory");
#pragma unroll 4
                for (int i = 1; i < nmma; ++i)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * (i & 3)), "l"(bd + 2 * (i & 3)), "r"(IDESC), "r"(1) : "memory");
            } else
            for (int i = 0; i < nmma; ++i) {
                const int kb = (i / 4) % kblocks_distinct, k = i % 4;
                const uint32_t a = s32(smem + kb * (16384 + N * 128));
                const uint64_t ad = make_desc(a) + 2 * k, bd = make_desc(a + 16384) + 2 * k;
                const uint32_t d = tmem + (uint32_t)(((i / 4) % nacc) * N);
                const uint32_t acc = i >= 4 * nacc || k != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(IDESC), "r"(acc) : "memory");
            }
            long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)), "r"(rep & 1) : "memory");
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
    }
    __syncthreads();
    if (threadIdx.x >= 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

template <int N> void run(long long* out) {
    const int smem = 4 * (16384 + N * 128) + 2048;
    cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nmma : {16, 64}) {
        rate<N><<<1, 64, smem>>>(nmma, 1, -1, out);
        cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        printf("N=%3d  LEAN loop       MMAs=%2d : issue %6lld clk (%5.1f / MMA), until commit completes %6lld clk (%5.1f / MMA)\n", N, nmma, h[0], (double)h[0] / nmma, h[1], (double)h[1] / nmma);
    }
    for (int nacc : {1, 4}) {
        if (nacc * N > 512) continue;
        for (int nmma : {16, 64}) {
            rate<N><<<1, 64, smem>>>(nmma, nacc, 4, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
            long long h[2];
            cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
            printf("N=%3d  accumulators=%d  MMAs=%2d : issue %6lld clk (%5.1f / MMA), until commit completes %6lld clk (%5.1f / MMA)\n", N, nacc, nmma, h[0], (double)h[0] / nmma, h[1],
                   (double)h[1] / nmma);
        }
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    run<32>(out); run<64>(out); run<128>(out); run<256>(out);
    return 0;
}
