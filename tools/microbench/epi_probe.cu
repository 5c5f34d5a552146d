// Epilogue probe for the tcgen05 render kernel (tools only): how fast can 8 warps of one CTA per SM move tensor-memory
// tiles to global memory (tcgen05.ld -> scale -> swizzled st.shared -> cp.async.bulk.tensor store), alone and next to
// warps that wait on an mbarrier (what the idle roles of the render kernel do)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench/_epi_probe tools/microbench/epi_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                \
    do {                                                                                     \
        cudaError_t e_ = (x);                                                                \
        if (e_ != cudaSuccess) {                                                             \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            exit(1);                                                                         \
        }                                                                                    \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32_issue(uint32_t addr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: TMA tensor store, `bufs` staged tiles per warp; mode 1: padded transpose + st.global.cs.v4
// spin: number of extra warps waiting on an mbarrier; passes: 1 or 2 reads of tensor memory per template
__global__ void __launch_bounds__(704, 1) epi_kernel(const __grid_constant__ CUtensorMap tmap, float *images, int n_tmpl, int mode,
                                                      int bufs, int spin, int passes, int hint, int n_ew, int *ticket) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ int s_t[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n_ew));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    if (warp < n_ew) {
        const int q = warp & 3, ch = warp >> 2, cstep = n_ew >> 2;
        const uint32_t tm_q = tm + ((uint32_t)(32 * q) << 16);
        const uint32_t tile_s = smem_u32(smem) + (uint32_t)warp * (uint32_t)bufs * 4096u;
        float *stg = reinterpret_cast<float *>(smem + (size_t)warp * 32 * 36 * 4);
        int nbuf = 0;
        float acc = 0.f;
        for (int it = 0;; ++it) {
            int t;
            if (ticket) {  // dynamic hand-out: one atomic per template, broadcast through shared memory
                if (threadIdx.x == 0) s_t[it & 1] = atomicAdd(ticket, 1);
                asm volatile("bar.sync 1, %0;" ::"r"(n_ew * 32) : "memory");
                t = s_t[it & 1];
            } else {
                t = blockIdx.x + it * gridDim.x;
            }
            if (t >= n_tmpl) break;
            if (passes == 2) {
                for (int h = 0; h < 2; ++h)
                    for (int ct = ch; ct < 8; ct += cstep) {
                        uint32_t r[32];
                        tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * ct), r);
                        tmem_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r[j]));
                    }
            }
            const float scale = acc == 123.f ? 2.f : 1.f;
            for (int h = 0; h < 2; ++h) {
                const int row0 = 128 * h + 32 * q;
                for (int ct = ch; ct < 8; ct += cstep) {
                    uint32_t r[32];
                    tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * ct), r);
                    tmem_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * scale;
                    if (mode == 0) {
                        if (lane == 0) {
                            if (bufs == 2)
                                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            else
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                        __syncwarp();
                        const uint32_t buf = tile_s + (uint32_t)(bufs == 2 ? (nbuf & 1) : 0) * 4096u;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4))),
                                         "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                         : "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap), "r"(buf),
                                         "r"(32 * ct), "r"(row0), "r"(t)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        ++nbuf;
                    } else {
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4 *>(stg + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        __syncwarp();
                        const int cg = lane & 7;
                        float *img = images + (size_t)t * 65536;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = 4 * i + (lane >> 3);
                            float4 o;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(smem_u32(stg + rr * 36 + 4 * cg)) : "memory");
                            __stcs(reinterpret_cast<float4 *>(img + (size_t)(row0 + rr) * 256 + 32 * ct + 4 * cg), o);
                        }
                    }
                }
            }
        }
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        if (acc == -5.f) images[0] = acc;
    } else if (warp < n_ew + spin) {
        uint32_t ok = 0;
        while (!ok) {
            if (hint)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0), "r"(0x989680u) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    const int n_tmpl = 16384;
    float *images;
    CK(cudaMalloc(&images, (size_t)n_tmpl * 65536 * 4));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    alignas(64) CUtensorMap tmap;
    const cuuint64_t dims[3] = {256, 256, (cuuint64_t)n_tmpl};
    const cuuint64_t strides[2] = {1024, 262144};
    const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
    CUresult cr = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, images, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)cr);
        return 1;
    }
    CK(cudaFuncSetAttribute(epi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    int *ticket;
    CK(cudaMalloc(&ticket, 4));
    struct V {
        const char *name;
        int mode, bufs, spin, passes, hint, n_ew, tick;
    } vs[] = {{"TMA 2 bufs, 2 passes, 8 warps, ticket", 0, 2, 0, 2, 0, 8, 1}, {"TMA 2 bufs, 2 passes, 16 warps, static", 0, 2, 0, 2, 0, 16, 0},
              {"TMA 2 bufs, 2 passes, 16 warps, ticket", 0, 2, 0, 2, 0, 16, 1}, {"TMA 2 bufs, 1 pass, 16 warps, ticket", 0, 2, 0, 1, 0, 16, 1},
              {"STG, 2 passes, 16 warps, ticket", 1, 2, 0, 2, 0, 16, 1}, {"STG, 2 passes, 8 warps, ticket", 1, 2, 0, 2, 0, 8, 1},{"TMA 2 bufs, 1 pass, alone", 0, 2, 0, 1, 0, 8, 0},      {"TMA 2 bufs, 2 passes, alone", 0, 2, 0, 2, 0, 8, 0},
              {"TMA 1 buf, 2 passes, alone", 0, 1, 0, 2, 0, 8, 0},     {"TMA 2 bufs, 2 passes, 6 waiters (no hint)", 0, 2, 6, 2, 0, 8, 0},
              {"TMA 2 bufs, 2 passes, 6 waiters (hint)", 0, 2, 6, 2, 1, 8, 0}, {"STG transpose, 1 pass, alone", 1, 2, 0, 1, 0, 8, 0},
              {"STG transpose, 2 passes, alone", 1, 2, 0, 2, 0, 8, 0}, {"STG transpose, 2 passes, 6 waiters (hint)", 1, 2, 6, 2, 1, 8, 0}};
    for (auto &v : vs) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(ticket, 0, 4));
            cudaEventRecord(e0);
            epi_kernel<<<148, 32 * (v.n_ew + v.spin), 140 * 1024>>>(tmap, images, n_tmpl, v.mode, v.bufs, v.spin, v.passes, v.hint, v.n_ew,
                                                                    v.tick ? ticket : nullptr);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-48s %8.3f ms  %7.1f GB/s\n", v.name, ms, (double)n_tmpl * 262144 / (ms * 1e-3) / 1e9);
    }
    return 0;
}
