// Micro-benchmark: sustained rate of legacy mma.sync.m16n8k16 (bf16 -> f32) on sm_100a, 16 independent
// accumulator tiles per warp (the K3 dense-path shape: a 32 x 64 warp region).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/microbench/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256, 2) k(float *out, int iters, uint32_t seed) {
    float c[16][4];
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a[2][4], b[8][2];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 4; ++j) a[i][j] = seed + threadIdx.x * 7 + i * 4 + j;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 2; ++j) b[i][j] = seed * 3 + threadIdx.x + i * 2 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) mma16816(c[m * 8 + n], a[m], b[n]);
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 2 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<148 * 2, 256>>>(out, iters, 0x3f803f80u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double mmas = 148.0 * 2 * 8 * iters * 16;
        printf("%.3f ms  %.1f G mma/s  %.1f TFLOP/s  (%.2f clk/mma/SM at 1.92 GHz)\n", ms, mmas / ms / 1e6,
               mmas * 4096 * 2 / ms / 1e9, ms * 1e-3 * 1.92e9 / (mmas / 148));
    }
    return 0;
}
