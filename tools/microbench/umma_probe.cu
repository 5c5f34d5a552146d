// tcgen05 probe for the dense K3 design (tools only; not part of the library).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench/_umma_probe tools/microbench/umma_probe.cu
//
// 1. correctness of hand-built shared-memory descriptors: D[128 x N] = A[128 x K] . B[K x N], bf16 in, float32
//    accumulators in TMEM, for the no-swizzle canonical layouts in both majors and both readings of LBO / SBO;
// 2. issue rate of back-to-back tcgen05.mma (M = 128, N = 128 / 256, K = 16) per SM, alone and while other warps
//    stream 16-byte stores into shared memory (what the operand producers of the render kernel do).
// Every wait is bounded: on a time-out the kernel sets a flag and returns, it never hangs the device.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                           \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                    \
        }                                                                               \
    } while (0)

__device__ int g_timeout;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t phase) {
    const long long t0 = clock64();
    while (!mbar_try(bar, phase)) {
        if (clock64() - t0 > 2000000000ll) {
            g_timeout = 1;
            return false;
        }
    }
    return true;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= 1ull << 46;  // descriptor version (Blackwell)
    return d;         // layout type 0: no swizzle
}
// kind::f16 instruction descriptor: float32 accumulate, bf16 x bf16
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define PROXY_FENCE() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")

// ---------------------------------------------------------------------------------------------------
// 1. correctness
// ---------------------------------------------------------------------------------------------------
// Element (mn, k) of an operand with MN rows, 8-row / 8-k core matrices of 128 contiguous bytes:
//   MN-major core: [k % 8][mn % 8]     K-major core: [mn % 8][k % 8]
// core (mn / 8, k / 8) at byte offset (mn / 8) * 128 + (k / 8) * (MN / 8) * 128.
__device__ __forceinline__ int elem_offset(int mn, int k, int MN, int mn_major) {
    const int core = (mn >> 3) * 64 + (k >> 3) * (MN >> 3) * 64;
    return core + (mn_major ? (k & 7) * 8 + (mn & 7) : (mn & 7) * 8 + (k & 7));
}

// variant bit 0: operands MN-major (1) or K-major (0); bit 1: swap the LBO / SBO fields
__global__ void __launch_bounds__(128) probe_correct(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int N, int K,
                                                     int variant) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int mn_major = variant & 1, swap = (variant >> 1) & 1;
    __nv_bfloat16 *sA = reinterpret_cast<__nv_bfloat16 *>(smem);
    __nv_bfloat16 *sB = sA + 128 * K;
    for (int e = threadIdx.x; e < 128 * K; e += blockDim.x) {
        const int m = e / K, k = e % K;
        sA[elem_offset(m, k, 128, mn_major)] = A[e];  // A[m][k]
    }
    for (int e = threadIdx.x; e < K * N; e += blockDim.x) {
        const int k = e / N, n = e % N;
        sB[elem_offset(n, k, N, mn_major)] = B[e];  // B[k][n]
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 256);
    PROXY_FENCE();
    TC_FENCE_BEFORE();
    __syncthreads();
    TC_FENCE_AFTER();
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(128, N, mn_major, mn_major);
        const uint32_t a_kgroup = 16 * 128, b_kgroup = (N / 8) * 128, mn_group = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            // MN-major, no swizzle (CUTLASS make_umma_desc): SBO = stride between MN groups, LBO = stride between
            // K groups.  K-major, no swizzle: SBO = stride between MN groups, LBO = stride between K groups as well.
            uint32_t a_lbo = a_kgroup, a_sbo = mn_group, b_lbo = b_kgroup, b_sbo = mn_group;
            if (swap) {
                a_lbo = mn_group, a_sbo = a_kgroup, b_lbo = mn_group, b_sbo = b_kgroup;
            }
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * a_kgroup, a_lbo, a_sbo);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * b_kgroup, b_lbo, b_sbo);
            umma(tm, da, db, idesc, ks > 0);
        }
        umma_commit(&bar);
    }
    const bool ok = mbar_wait_bounded(&bar, 0);
    TC_FENCE_AFTER();
    if (ok) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c = 0; c < N; c += 32) {
            uint32_t v[32];
            tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
        }
    }
    TC_FENCE_BEFORE();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(tm, 256);
}

// ---------------------------------------------------------------------------------------------------
// 2. issue rate
// ---------------------------------------------------------------------------------------------------
// warp 0 lane 0 issues `iters` MMAs (three per 16-spot stage, cycling over `stages` operand buffers); warps
// 1 .. n_noise stream STS.128 into a scratch area for the whole time.
__global__ void __launch_bounds__(288) probe_rate(int N, int iters, int stages, int n_noise, long long *cycles_out,
                                                  float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage_bytes = 2 * (128 * 16 * 2) + 2 * (N * 16 * 2);  // A hi, A lo, B hi, B lo for 16 spots
    for (int e = threadIdx.x; e < (int)(stage_bytes * stages / 4); e += blockDim.x)
        reinterpret_cast<uint32_t *>(smem)[e] = 0x3c003c00u + e;  // some finite bf16 values
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    PROXY_FENCE();
    TC_FENCE_BEFORE();
    __syncthreads();
    TC_FENCE_AFTER();
    const uint32_t tm = tmem_base;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, N, 1, 1);
            const uint32_t a_bytes = 128 * 16 * 2, b_bytes = N * 16 * 2;
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t st = smem_u32(smem) + (uint32_t)(i % stages) * stage_bytes;
                const uint64_t a_hi = make_desc(st, 16 * 128, 128), a_lo = make_desc(st + a_bytes, 16 * 128, 128);
                const uint64_t b_hi = make_desc(st + 2 * a_bytes, (N / 8) * 128, 128);
                const uint64_t b_lo = make_desc(st + 2 * a_bytes + b_bytes, (N / 8) * 128, 128);
                const uint32_t d = tm + (uint32_t)((i & 1) * N);
                umma(d, a_hi, b_hi, idesc, 1);
                umma(d, a_hi, b_lo, idesc, 1);
                umma(d, a_lo, b_hi, idesc, 1);
            }
            umma_commit(&bar);
            mbar_wait_bounded(&bar, 0);
            const long long t1 = clock64();
            cycles_out[blockIdx.x] = t1 - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp <= n_noise) {
        // 16-byte stores, conflict-free (a quarter warp covers 128 contiguous bytes), into the area behind the stages
        uint32_t base = smem_u32(smem) + stage_bytes * stages + (uint32_t)(warp - 1) * 4096 + lane * 16;
        float acc = 0.f;
        int n = 0;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(base + (uint32_t)(j * 512)), "f"(acc) : "memory");
            acc += 1.f;
            if (++n > 50000000) break;
        }
        if (acc == -1.f) sink[0] = acc;
    }
    TC_FENCE_BEFORE();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
    int zero = 0;
    CK(cudaMemcpyToSymbol(g_timeout, &zero, sizeof(int)));

    // ---- 1. correctness
    for (int N : {128, 256}) {
        const int K = 32;
        std::vector<float> A(128 * K), B(K * N), ref(128 * N, 0.f);
        srand(1);
        for (auto &x : A) x = bf16_round((rand() % 2001 - 1000) / 500.f);
        for (auto &x : B) x = bf16_round((rand() % 2001 - 1000) / 500.f);
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[k * N + n];
                ref[m * N + n] = (float)s;
            }
        std::vector<__nv_bfloat16> Ah(A.size()), Bh(B.size());
        for (size_t i = 0; i < A.size(); ++i) Ah[i] = __float2bfloat16(A[i]);
        for (size_t i = 0; i < B.size(); ++i) Bh[i] = __float2bfloat16(B[i]);
        __nv_bfloat16 *dA, *dB;
        float *dD;
        CK(cudaMalloc(&dA, Ah.size() * 2));
        CK(cudaMalloc(&dB, Bh.size() * 2));
        CK(cudaMalloc(&dD, ref.size() * 4));
        CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
        for (int variant = 0; variant < 2; ++variant) {  // (the swapped readings fault: measured)
            CK(cudaMemset(dD, 0xff, ref.size() * 4));
            const size_t smem = (size_t)(128 + N) * K * 2;
            probe_correct<<<1, 128, smem>>>(dA, dB, dD, N, K, variant);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("correct N=%d variant=%d: CUDA error %s\n", N, variant, cudaGetErrorString(e));
                return 2;  // sticky error: nothing else can run
            }
            std::vector<float> got(ref.size());
            CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
            double worst = 0;
            for (size_t i = 0; i < got.size(); ++i) {
                double d = fabs((double)got[i] - ref[i]);
                if (!(d == d)) d = 1e30;
                worst = fmax(worst, d);
            }
            int to = 0;
            CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int)));
            printf("correct N=%3d %s-major lbo/sbo %s: max abs err %.3g%s\n", N, (variant & 1) ? "MN" : "K ",
                   (variant & 2) ? "swapped " : "as-read", worst, to ? "  (TIMEOUT)" : "");
            CK(cudaMemcpyToSymbol(g_timeout, &zero, sizeof(int)));
        }
        cudaFree(dA), cudaFree(dB), cudaFree(dD);
    }

    // ---- 2. rate
    long long *d_cycles;
    float *d_sink;
    CK(cudaMalloc(&d_cycles, sizeof(long long) * 1024));
    CK(cudaMalloc(&d_sink, 64));
    CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int sms = prop.multiProcessorCount;
    for (int grid : {1, sms}) {
        for (int N : {128, 256}) {
            for (int n_noise : {0, 4, 8}) {
                const int iters = 4000, stages = 4;
                const size_t smem = (size_t)stages * (2 * 4096 + 2 * N * 32) + 8 * 4096 + 1024;
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0), cudaEventCreate(&e1);
                probe_rate<<<grid, 288, smem>>>(N, 100, stages, n_noise, d_cycles, d_sink);  // warm-up
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0);
                probe_rate<<<grid, 288, smem>>>(N, iters, stages, n_noise, d_cycles, d_sink);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("rate: CUDA error %s\n", cudaGetErrorString(e));
                    return 2;
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                std::vector<long long> cyc(grid);
                CK(cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (auto c : cyc) mx = c > mx ? c : mx;
                const double per_mma = (double)mx / (3.0 * iters);
                const double flops = 2.0 * 128 * N * 16 * 3.0 * iters * grid;
                printf("rate grid=%3d N=%3d noise_warps=%d: %.1f cycles per MMA (ideal %d), %.1f TFLOP/s over the launch (%.3f ms)\n",
                       grid, N, n_noise, per_mma, N / 2, flops / (ms * 1e-3) / 1e12, ms);
            }
        }
    }
    int to = 0;
    CK(cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int)));
    if (to) printf("a wait timed out\n");
    return 0;
}
