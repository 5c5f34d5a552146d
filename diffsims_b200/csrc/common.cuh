// Shared helpers for the diffsims_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/diffsims_b200.h"

namespace ds {

void set_error(const char *fmt, ...);
int num_sms();  // of the current device (queried per call: one process may drive several GPUs)

// Process-wide tuning options (cabi.cu): read from the environment once, changed with ds_set_option.
// -1 = not set.
enum Opt {
    OPT_RENDER_PIPE = 0,
    OPT_RENDER_GROUP,
    OPT_RENDER_FRONTS,
    OPT_RENDER_PIPE_MAXCAP,
    OPT_RENDER_NOSTAGE,
    OPT_RENDER_MMA,
    OPT_RENDER_MMA_MIN,
    OPT_RENDER_MMA_TMPL_MIN,
    OPT_RENDER_UMMA,
    OPT_RENDER_UMMA_WINDOW,
    OPT_RENDER_ZERO_TMA,
    OPT_RENDER_UMMA_TEAM,
    OPT_RENDER_ROWS,
    OPT_RENDER_ROWS_STAGES,
    OPT_SIM_LINES,
    OPT_SIM_SPLIT,
    OPT_SIM_CTA,
    OPT_SIM_STASH,
    OPT_COUNT
};
int option(Opt o);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

#define DS_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            ds::set_error(__VA_ARGS__); \
            return -1;                 \
        }                              \
    } while (0)

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
// Waits are bounded: a barrier that does not complete within ~2^28 polls (each try_wait suspends in hardware for a
// while, so this is many seconds) is a protocol bug, and the kernel traps -- the launch fails loudly with a sticky
// error instead of hanging the device.  The bound is a poll counter, not a clock read: idle roles of the
// warp-specialised kernels sit in this loop and every instruction in it competes with the working warps.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "add.u32 n, n, 1;\n"
        "setp.gt.u32 q, n, 0x10000000;\n"
        "@q trap;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// one elected lane of a converged warp (elect.sync): the compiler treats the predicate as warp-uniform, so uniform-
// datapath instructions (tcgen05.mma, cp.async.bulk.tensor, ...) issue without a per-lane loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace ds
