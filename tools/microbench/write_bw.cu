// Micro-benchmark: pure-write HBM bandwidth of template-shaped store patterns on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/write_bw tools/microbench/write_bw.cu && /tmp/write_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int FLAVOR>
__device__ __forceinline__ void st4(float4 *p, float4 v) {
    if (FLAVOR == 0) *p = v;
    else if (FLAVOR == 1) __stcs(p, v);
    else if (FLAVOR == 2) __stcg(p, v);
    else __stwt(p, v);
}

// mode 0: CTA sweeps its template linearly (4 KB per step)
// mode 1: K3 pattern: warp region 64 x 32 px, lane tile 8 x 8 as two float4 column groups 32 px apart
// mode 2: warp writes full 1 KB rows: region 256 px x 8 rows, lane = 8 consecutive px (2 float4), rows strided
// mode 3: like 1 but each lane's 8 px contiguous (32 B), region 256 x 8? no: region 64 x 32, lane 8 px contiguous
__device__ int g_ticket;
template <int FLAVOR>
__global__ void __launch_bounds__(256, 2) write_kernel(float *img, int n_tmpl, int mode, float val) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int s_t;
    if (mode == 6) {  // dynamic template assignment through a global ticket
        const int lx = lane & 7, ly = lane >> 3;
        while (true) {
            __syncthreads();
            if (threadIdx.x == 0) s_t = atomicAdd(&g_ticket, 1);
            __syncthreads();
            const int t = s_t;
            if (t >= n_tmpl) return;
            float *base = img + (size_t)t * 65536;
            const float4 v = make_float4(val + t, val, val, val);
            for (int reg = warp; reg < 32; reg += 8) {
                const int rx0 = (reg & 3) * 64, ry0 = (reg >> 2) * 32;
                float *dst = base + (size_t)(ry0 + 8 * ly) * 256 + rx0 + 4 * lx;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256), v);
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256 + 32), v);
                }
            }
        }
    }
    const int per = (n_tmpl + gridDim.x - 1) / gridDim.x;
    const int t_begin = (mode == 5) ? blockIdx.x * per : blockIdx.x;
    const int t_end = (mode == 5) ? min(n_tmpl, t_begin + per) : n_tmpl;
    const int t_step = (mode == 5) ? 1 : gridDim.x;
    for (int t = t_begin; t < t_end; t += t_step) {
        float *base = img + (size_t)t * 65536;
        const float4 v = make_float4(val + t, val, val, val);
        if (mode == 0) {
            for (int i = threadIdx.x; i < 16384; i += 256) st4<FLAVOR>(reinterpret_cast<float4 *>(base) + i, v);
        } else if (mode == 1) {
            const int lx = lane & 7, ly = lane >> 3;
            for (int reg = warp; reg < 32; reg += 8) {
                const int rx0 = (reg & 3) * 64, ry0 = (reg >> 2) * 32;
                float *dst = base + (size_t)(ry0 + 8 * ly) * 256 + rx0 + 4 * lx;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256), v);
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256 + 32), v);
                }
            }
        } else if (mode == 2) {
            for (int band = warp; band < 32; band += 8) {  // 8 rows per band
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float *dst = base + (size_t)(band * 8 + i) * 256 + lane * 8;
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst), v);
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + 4), v);
                }
            }
        } else if (mode == 4) {  // mode 1 with the region order rotated per CTA (decorrelates the offsets)
            const int lx = lane & 7, ly = lane >> 3;
            const int rot = (blockIdx.x * 5) & 31;
            for (int r0 = warp; r0 < 32; r0 += 8) {
                const int reg = (r0 + rot) & 31;
                const int rx0 = (reg & 3) * 64, ry0 = (reg >> 2) * 32;
                float *dst = base + (size_t)(ry0 + 8 * ly) * 256 + rx0 + 4 * lx;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256), v);
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256 + 32), v);
                }
            }
        } else if (mode == 5) {  // mode 1, templates assigned in contiguous chunks per CTA instead of strided
            const int lx = lane & 7, ly = lane >> 3;
            for (int reg = warp; reg < 32; reg += 8) {
                const int rx0 = (reg & 3) * 64, ry0 = (reg >> 2) * 32;
                float *dst = base + (size_t)(ry0 + 8 * ly) * 256 + rx0 + 4 * lx;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256), v);
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256 + 32), v);
                }
            }
        } else {
            const int lx = lane & 7, ly = lane >> 3;
            for (int reg = warp; reg < 32; reg += 8) {
                const int rx0 = (reg & 3) * 64, ry0 = (reg >> 2) * 32;
                float *dst = base + (size_t)(ry0 + 8 * ly) * 256 + rx0 + 8 * lx;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256), v);
                    st4<FLAVOR>(reinterpret_cast<float4 *>(dst + i * 256 + 4), v);
                }
            }
        }
    }
}

int main() {
    const int n = 32768;
    float *img;
    cudaMalloc(&img, (size_t)n * 65536 * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const char *fl[] = {"default", "cs", "cg", "wt"};
    for (int grid : {296, 592, 32768}) {
        for (int mode : {1, 4, 5, 6}) {
            for (int f = 1; f < 2; ++f) {
                float best = 1e9;
                for (int rep = 0; rep < 4; ++rep) {
                    int zero = 0;
                    cudaMemcpyToSymbol(g_ticket, &zero, sizeof(int));
                    cudaEventRecord(a);
                    if (f == 0) write_kernel<0><<<grid, 256>>>(img, n, mode, 1.f);
                    if (f == 1) write_kernel<1><<<grid, 256>>>(img, n, mode, 1.f);
                    if (f == 2) write_kernel<2><<<grid, 256>>>(img, n, mode, 1.f);
                    if (f == 3) write_kernel<3><<<grid, 256>>>(img, n, mode, 1.f);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    float ms;
                    cudaEventElapsedTime(&ms, a, b);
                    if (rep > 0 && ms < best) best = ms;
                }
                printf("grid %5d mode %d %-7s %8.1f us %7.0f GB/s\n", grid, mode, fl[f], best * 1e3,
                       (double)n * 262144 / best / 1e6);
            }
        }
    }
    // cudaMemset for reference
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        cudaMemsetAsync(img, 0, (size_t)n * 65536 * 4);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("cudaMemset %8.1f us %7.0f GB/s\n", ms * 1e3, (double)n * 262144 / ms / 1e6);
    }
    return 0;
}
