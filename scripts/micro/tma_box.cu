// Microbenchmark (debug aid): how fast do the CTAs of a persistent grid pull the SAME [rows x 1024] bf16 activation block from L2
// with TMA, as a function of the box shape?  Mirrors the A-operand fetch of the weights-stationary recurrent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_box tma_box.cu -lcuda && ./tma_box
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
}

// mode 0: 2-D boxes {64 cols, rows}, one op per K-block (16 ops).  mode 1: 3-D boxes {64, rows, kb_per_op}.
__global__ void __launch_bounds__(128) pull(const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, int mode, int rows, int kb_per_op,
                                            int iters, int distinct, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const int row0 = distinct ? blockIdx.x * rows : 0;
    long long t0 = 0, total = 0;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            t0 = clock64();
            const int nops = 16 / kb_per_op;
            for (int i = 0; i < nops; ++i) {
                const uint32_t bytes = (uint32_t)rows * 128u * kb_per_op;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + i)), "r"(bytes) : "memory");
                if (mode == 0)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(s32(smem + i * rows * 128)), "l"(&map2), "r"(s32(bar + i)), "r"(i * 64), "r"(row0) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(s32(smem + i * kb_per_op * rows * 128)), "l"(&map3), "r"(s32(bar + i)), "r"(0), "r"(row0), "r"(i * kb_per_op) : "memory");
            }
            for (int i = 0; i < nops; ++i) mbar_wait(bar + i, it & 1);
            total += clock64() - t0;
        }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = total;
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int K = 1024, ROWS_TOTAL = 64 * 148;
    __nv_bfloat16* A;
    CK(cudaMalloc(&A, (size_t)ROWS_TOTAL * K * 2));
    CK(cudaMemset(A, 0, (size_t)ROWS_TOTAL * K * 2));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    Enc enc = (Enc)fp;
    long long* out; CK(cudaMalloc(&out, 148 * 8));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("SM clock %d kHz\n", clk_khz);
    printf("rows  ops x KB   same-region: clk (us)   GB/s per SM | distinct regions: clk (us)  GB/s per SM   [grid 128]\n");
    for (int rows : {64, 128}) {
        for (int kb : {1, 2, 4, 8, 16}) {
            if (rows * 128 * kb > 200 * 1024 / (16 / kb) * 16 / kb && rows * 128 * 16 > 200 * 1024) continue;
            CUtensorMap m2, m3;
            cuuint64_t d2[2] = {(cuuint64_t)K, (cuuint64_t)ROWS_TOTAL}; cuuint64_t s2[1] = {(cuuint64_t)K * 2};
            cuuint32_t b2[2] = {64, (cuuint32_t)rows}, e2[2] = {1, 1};
            if (enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 2d failed\n"); return 1; }
            cuuint64_t d3[3] = {64, (cuuint64_t)ROWS_TOTAL, (cuuint64_t)(K / 64)}; cuuint64_t s3[2] = {(cuuint64_t)K * 2, 128};
            cuuint32_t b3[3] = {64, (cuuint32_t)rows, (cuuint32_t)kb}, e3[3] = {1, 1, 1};
            if (enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 3d failed (kb %d)\n", kb); continue; }
            const int smem = rows * 128 * 16 + 2048;
            CK(cudaFuncSetAttribute(pull, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const int iters = 200;
            double res[2];
            for (int distinct = 0; distinct < 2; ++distinct) {
                const int mode = kb == 1 ? 0 : 1;
                pull<<<128, 128, smem>>>(m2, m3, mode, rows, kb, 20, distinct, out);
                pull<<<128, 128, smem>>>(m2, m3, mode, rows, kb, iters, distinct, out);
                CK(cudaDeviceSynchronize());
                long long h[128]; CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
                double mx = 0; for (int i = 0; i < 128; ++i) mx = h[i] > mx ? (double)h[i] : mx;
                res[distinct] = mx / iters;
            }
            const double bytes = rows * 128.0 * 16;
            printf("%4d  %2d x %3d   %8.0f (%5.2f)  %7.1f | %8.0f (%5.2f)  %7.1f\n", rows, 16 / kb, rows * 128 * kb / 1024, res[0], res[0] / clk_khz * 1e3, bytes / (res[0] / clk_khz * 1e-3) / 1e9,
                   res[1], res[1] / clk_khz * 1e3, bytes / (res[1] / clk_khz * 1e-3) / 1e9);
        }
    }
    return 0;
}
