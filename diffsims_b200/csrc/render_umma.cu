// K3 on the 5th-generation tensor cores: templates of up to 256 x 256 pixels as per-template matrix products.
//
// The blurred image of a spot list is the rank-S product  out[y][x] = sum_s (a_s Wy_s[y]) Wx_s[x]  with the
// reflect-folded 1-D Gaussian weights Wy_s / Wx_s of spot s (render.cu explains the folding).  Here that product
// runs on tcgen05.mma with the accumulators in tensor memory:
//
//   * a template is two HALVES of 128 rows (UMMA M = 128); a half only sums over the spots whose box reaches it;
//   * spots are taken 16 at a time (one UMMA K step).  A producer warp writes, ONCE per (half, chunk), the
//     operand tiles  A = a_s Wy_s[y]  (128 x 16)  and  B = Wx_s[x]  (16 x N)  to shared memory as bf16 high and
//     low parts in the canonical MN-major no-swizzle layout (8 x 8 core matrices of 128 contiguous bytes), and
//     one elected lane of the MMA warp issues three products  A_hi B_hi + A_hi B_lo + A_lo B_hi  (float32
//     accumulation, ~16 mantissa bits: < 3e-5 of the peak, the parity bound is 1e-4);
//   * after the first chunk of a half (which covers, and thereby zeroes, all columns) a chunk only spans the
//     16-column-aligned window its spots reach: N is a per-instruction field, and the front warp sorts each
//     half's spots by column so that the windows of a dense template are narrow;
//   * four epilogue warps read the accumulators back with tcgen05.ld (lane = image row), take the template
//     maximum from them (no second evaluation), scale, transpose 32 x 32 tiles through padded shared memory and
//     stream them out with the same 4 rows x 128 bytes st.global.cs.v4 pattern as the other render kernels.
//
// Warp roles of the persistent CTA (one per SM, templates drawn from the global ticket):
//   warp 0        front: prefetched spot rows -> float64 projection -> last-write-wins hash -> compaction ->
//                 per-half lists (sorted by column) -> publishes one of two template slots
//   warp 1        MMA issue (lane 0), owns the tensor-memory allocation (512 columns = 2 halves x 256)
//   warps 2..5    epilogue (tensor-memory lane quarter = warp % 4)
//   warps 6..11   operand producers, one shared-memory stage each
// mbarriers: slot full/empty (front <-> everyone), stage full/empty (producer <-> tcgen05.commit),
// half full/empty (tcgen05.commit <-> epilogue).  Bound: tensor pipe for dense templates, HBM write otherwise.
//
// Reference: diffsims/pattern/detector_functions.py:293-300 (assignment + scipy.ndimage.gaussian_filter),
// diffsims/simulations/simulation2d.py:261-285, :422-441.
#include "render_device.cuh"

namespace ds {

constexpr int UM_NP = 6;                       // producer warps = operand stages
constexpr int UM_EPI = 4;                      // epilogue warps
constexpr int UM_WARPS = 2 + UM_EPI + UM_NP;   // 12
constexpr int UM_THREADS = UM_WARPS * 32;      // 384
constexpr int UM_A_BYTES = 128 * 16 * 2;       // one of A_hi / A_lo:  16 MN groups x 2 K groups x 128 B
constexpr int UM_B_BYTES = 256 * 16 * 2;       // one of B_hi / B_lo:  32 MN groups x 2 K groups x 128 B
constexpr int UM_STAGE_BYTES = 2 * UM_A_BYTES + 2 * UM_B_BYTES;  // 24 KB
constexpr int UM_EPI_PITCH = 36;               // floats per staged row (32 + 4: conflict-free both ways)
constexpr int UM_EPI_BYTES = 32 * UM_EPI_PITCH * 4;
constexpr int UM_BPAD = 8;                     // zero padding of the bf16 tap table on either side
constexpr int UM_MAX_CAP = 1024;

struct UmHeader {  // 32 bytes at the start of a slot
    int t, n_live, n_half[2], pad[4];
};

// entries (16 bytes = 8 taps) per shifted copy of the bf16 tap table
__host__ __device__ inline int um_n8(int radius) { return (2 * radius + 2 * UM_BPAD + 1 + 7) / 8 + 1; }

// ---- tcgen05 wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // shared-memory matrix descriptor, no swizzle: start address, leading (K-group) and stride (MN-group) byte
    // offsets in 16-byte units, descriptor version 1 (validated by tools/microbench/umma_probe.cu)
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc(int n_cols) {
    // kind::f16: float32 accumulators, bf16 x bf16, both operands MN-major, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
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
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// volatile loads for shared memory that other warps (or this warp, through st.shared) rewrite between reads
__device__ __forceinline__ float4 lds128v(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64v(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// One 16-byte operand unit (8 consecutive pixels of one spot): split into bf16 high / low parts and store.
__device__ __forceinline__ void store_split8(uint32_t hi_addr, uint32_t lo_addr, float4 w0, float4 w1, float a) {
    uint32_t h[4], l[4];
    bf16_split2(a * w0.x, a * w0.y, h[0], l[0]);
    bf16_split2(a * w0.z, a * w0.w, h[1], l[1]);
    bf16_split2(a * w1.x, a * w1.y, h[2], l[2]);
    bf16_split2(a * w1.z, a * w1.w, h[3], l[3]);
    sts128(hi_addr, h[0], h[1], h[2], h[3]);
    sts128(lo_addr, l[0], l[1], l[2], l[3]);
}

template <bool VEC>
__global__ void __launch_bounds__(UM_THREADS, 1) render_umma_kernel(const RenderParams p, const int slot_bytes,
                                                                     const int front_bytes, const int window) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_slot_full[2], s_slot_empty[2], s_stage_full[UM_NP], s_stage_empty[UM_NP],
        s_half_full[2], s_half_empty[2], s_rows[2];
    __shared__ int2 s_chunk[UM_NP];   // (first column, columns) of the chunk in each stage
    __shared__ float s_emax[2][UM_EPI];
    __shared__ uint32_t s_tmem;
    __shared__ double s_norm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.radius, H = p.H, W = p.W;
    const int n_halves = (H + 127) >> 7;
    const int Wp = (W + 15) & ~15;  // columns of a full-width product
    const int n8 = um_n8(R);

    // ---- shared memory: stages | epilogue staging | tap LUT (float, 4 shifted copies) | bf16 tap table (8 shifted
    // copies, hi then lo) | front workspace | 2 slots --------------------------------------------------------------
    unsigned char *stages = smem_raw;
    unsigned char *epi = stages + (size_t)UM_NP * UM_STAGE_BYTES;
    float4 *lut = reinterpret_cast<float4 *>(epi + (size_t)UM_EPI * UM_EPI_BYTES);
    unsigned char *btab = reinterpret_cast<unsigned char *>(lut) + lut_smem_bytes(p.n4);
    unsigned char *front_ws = btab + (size_t)2 * 8 * n8 * 16;
    unsigned char *slots = front_ws + front_bytes;
    auto slot_header = [&](int s) { return reinterpret_cast<UmHeader *>(slots + (size_t)s * slot_bytes); };
    auto slot_spots = [&](int s) { return reinterpret_cast<uint2 *>(slots + (size_t)s * slot_bytes + 32); };
    auto slot_list = [&](int s, int h) {
        return reinterpret_cast<unsigned short *>(slots + (size_t)s * slot_bytes + 32 + (size_t)p.cap * 8) + (size_t)h * p.cap;
    };

    // ---- CTA-wide, once ---------------------------------------------------------------------------------------
    if (warp == 0) {
        double part = 0.0;
        for (int k = lane; k <= R; k += 32) part += (k == 0 ? 1.0 : 2.0) * exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
            s_norm = part;
            for (int s = 0; s < 2; ++s) {
                mbar_init(&s_slot_full[s], 1);
                mbar_init(&s_slot_empty[s], 1 + UM_EPI + UM_NP);
                mbar_init(&s_half_full[s], 1);
                mbar_init(&s_half_empty[s], UM_EPI);
                mbar_init(&s_rows[s], 1);
            }
            for (int s = 0; s < UM_NP; ++s) {
                mbar_init(&s_stage_full[s], 1);
                mbar_init(&s_stage_empty[s], 1);
            }
            fence_mbar_init();
        }
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    fill_lut(lut, p.n4, R, p.sigma, 1.0 / s_norm, threadIdx.x, UM_THREADS);
    {   // bf16 table: copy sh (0..7), entry i holds taps L[8 i + sh .. + 7] of the padded kernel L[a] = w[|a - (R + BPAD)|]
        const double inv_norm = 1.0 / s_norm;
        unsigned short *bh = reinterpret_cast<unsigned short *>(btab);
        unsigned short *bl = bh + (size_t)8 * n8 * 8;
        for (int e = threadIdx.x; e < 8 * n8 * 8; e += UM_THREADS) {
            const int sh = e / (n8 * 8), rem = e % (n8 * 8);
            const int k = abs(rem + sh - (R + UM_BPAD));
            const float w = (k <= R) ? (float)(exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k) * inv_norm) : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            bh[e] = __bfloat16_as_ushort(hi);
            bl[e] = __bfloat16_as_ushort(__float2bfloat16_rn(w - __bfloat162float(hi)));
        }
    }
    __syncthreads();
    const uint32_t tm = s_tmem;

    if (warp == 0) {
        // =============================== front warp ==============================================================
        unsigned char *base = front_ws;
        double *stage_rows[2] = {nullptr, nullptr};
        if (p.stage) {
            stage_rows[0] = reinterpret_cast<double *>(base);
            stage_rows[1] = stage_rows[0] + p.cap * 4;
            base += (size_t)2 * p.cap * 32;
        }
        unsigned long long *hash = reinterpret_cast<unsigned long long *>(base);
        base += (size_t)p.table_size * 8;
        int *key = reinterpret_cast<int *>(base);
        base += (size_t)p.cap * 4;
        float *inten = reinterpret_cast<float *>(base);
        base += (size_t)p.cap * 4;
        int *bins = reinterpret_cast<int *>(base);  // [W + 1] counting sort by column
        auto prefetch = [&](int t, int buf) {
            const uint32_t bx = (uint32_t)p.cap * 24u, bi = (uint32_t)p.cap * 8u;
            mbar_expect_tx(&s_rows[buf], bx + bi);
            bulk_g2s(stage_rows[buf], p.xyz + (size_t)t * p.cap * 3, bx, &s_rows[buf]);
            bulk_g2s(stage_rows[buf] + p.cap * 3, p.intensity + (size_t)t * p.cap, bi, &s_rows[buf]);
        };
        auto draw = [&]() {
            int t = 0;
            if (lane == 0) t = atomicAdd(&p.ticket[0], 1);
            return __shfl_sync(0xffffffffu, t, 0);
        };
        int t = draw();
        int n_next = 0;
        if (t < p.n_tmpl) {
            n_next = p.count[t];
            if (p.stage && lane == 0) prefetch(t, 0);
        }
        for (int k = 0;; ++k) {
            const int slot = k & 1, buf = k & 1;
            mbar_wait(&s_slot_empty[slot], ((uint32_t)(k >> 1) & 1u) ^ 1u);
            UmHeader *hd = slot_header(slot);
            if (t >= p.n_tmpl) {  // out of work: stop slot
                if (lane == 0) {
                    hd->t = -1;
                    mbar_arrive(&s_slot_full[slot]);
                }
                break;
            }
            const int n = min(n_next, p.cap);
            const int t_next = draw();
            if (t_next < p.n_tmpl) {
                n_next = p.count[t_next];
                if (p.stage && lane == 0) prefetch(t_next, buf ^ 1);
            }
            const double *sxyz = p.xyz + (size_t)t * p.cap * 3;
            const double *sint = p.intensity + (size_t)t * p.cap;
            if (p.stage) {
                mbar_wait(&s_rows[buf], (uint32_t)(k >> 1) & 1u);
                sxyz = stage_rows[buf];
                sint = stage_rows[buf] + p.cap * 3;
            }
            uint2 *spots = slot_spots(slot);

            // ---- project, last-write-wins, compact (simulation2d.py:261-285, :422-430; detector_functions.py:297)
            for (int e = lane; e < p.table_size; e += 32) hash[e] = 0ull;
            for (int e = lane; e <= W; e += 32) bins[e] = 0;
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                const double xs = sxyz[3 * j] / p.cal, ys = sxyz[3 * j + 1] / p.cal;
                const double px = xs * p.ca - p.mirror * ys * p.sa + p.cx;
                const double py = p.mirror * ys * p.ca + xs * p.sa + p.cy;
                int kk = -1;
                if (px >= 0.0 && px < (double)W && py >= 0.0 && py < (double)H) {
                    kk = (int)py * W + (int)px;
                    const unsigned long long packed = ((unsigned long long)(kk + 1) << 32) | (unsigned)j;
                    unsigned h = ((unsigned)kk * 2654435761u) & (p.table_size - 1);
                    while (true) {
                        unsigned long long cur = hash[h];
                        if (cur == 0ull) {
                            const unsigned long long old = atomicCAS(&hash[h], 0ull, packed);
                            if (old == 0ull) break;
                            cur = old;
                        }
                        if ((cur >> 32) == (unsigned long long)(kk + 1)) {
                            atomicMax(&hash[h], packed);
                            break;
                        }
                        h = (h + 1) & (p.table_size - 1);
                    }
                }
                key[j] = kk;
                inten[j] = (float)sint[j];
            }
            __syncwarp();
            // live spots in list order; with the column windows on, a stable counting sort by column decides their
            // place (narrow windows per chunk of 16; stable = the float32 sums do not depend on scheduling)
            int n_live = 0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                bool live = false;
                int kk = -1;
                if (j < n && (kk = key[j]) >= 0) {
                    unsigned h = ((unsigned)kk * 2654435761u) & (p.table_size - 1);
                    while ((hash[h] >> 32) != (unsigned long long)(kk + 1)) h = (h + 1) & (p.table_size - 1);
                    live = (unsigned)(hash[h] & 0xffffffffu) == (unsigned)j;
                }
                if (j < n && !live) key[j] = -1;
                if (live && window) atomicAdd(&bins[kk % W + 1], 1);
                const unsigned mask = __ballot_sync(0xffffffffu, live);
                if (live && !window) {
                    const int d = n_live + __popc(mask & ((1u << lane) - 1u));
                    spots[d] = make_uint2((unsigned)(kk % W) | ((unsigned)(kk / W) << 16), __float_as_uint(inten[j]));
                }
                n_live += __popc(mask);
            }
            __syncwarp();
            if (window) {
                // exclusive scan of the column histogram (bins[c + 1] = spots in column c) ...
                int carry = 0;
                for (int c0 = 0; c0 <= W; c0 += 32) {
                    const int c = c0 + lane;
                    int incl = c <= W ? bins[c] : 0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    if (c <= W) bins[c] = carry + incl;  // = first slot of column c
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                __syncwarp();
                // ... then a stable scatter: 32 spots at a time, lanes of one column ranked by lane number
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    const int kk = j < n ? key[j] : -1;
                    const unsigned mask = __ballot_sync(0xffffffffu, kk >= 0);
                    int col = 0, slot_at = 0;
                    unsigned peers = 0;
                    if (kk >= 0) {
                        col = kk % W;
                        peers = __match_any_sync(mask, col);
                        slot_at = bins[col] + __popc(peers & ((1u << lane) - 1u));
                    }
                    __syncwarp();
                    if (kk >= 0) {
                        spots[slot_at] = make_uint2((unsigned)col | ((unsigned)(kk / W) << 16), __float_as_uint(inten[j]));
                        if ((peers & ((1u << lane) - 1u)) == 0u) bins[col] += __popc(peers);
                    }
                    __syncwarp();
                }
            }
            // ---- per-half lists: the spots whose box reaches rows [128 h, 128 h + 127] (the folded images of a
            // spot lie inside its own clipped box)
            int n_half[2] = {0, 0};
            for (int h = 0; h < n_halves; ++h) {
                unsigned short *list = slot_list(slot, h);
                int cnt = 0;
                for (int j0 = 0; j0 < n_live; j0 += 32) {
                    const int j = j0 + lane;
                    bool hit = false;
                    if (j < n_live) {
                        const int sy = (int)(spots[j].x >> 16);
                        hit = sy + R >= 128 * h && sy - R <= 128 * h + 127;
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, hit);
                    if (hit) list[cnt + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)j;
                    cnt += __popc(mask);
                }
                n_half[h] = cnt;
            }
            if (lane == 0) {
                hd->t = t;
                hd->n_live = n_live;
                hd->n_half[0] = n_half[0];
                hd->n_half[1] = n_half[1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_full[slot]);
            t = t_next;
        }
    } else if (warp == 1) {
        // =============================== MMA issue ===============================================================
        int c = 0;  // running chunk number: chunk c lives in stage c % UM_NP
        for (int k = 0;; ++k) {
            const int slot = k & 1;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k >> 1) & 1u);
            const UmHeader *hd = slot_header(slot);
            if (hd->t < 0) break;
            for (int h = 0; h < n_halves; ++h) {
                const int n_chunks = max(1, (hd->n_half[h] + 15) >> 4);
                mbar_wait(&s_half_empty[h], ((uint32_t)k & 1u) ^ 1u);  // the epilogue has drained template k - 1
                tc_fence_after();
                for (int j = 0; j < n_chunks; ++j, ++c) {
                    const int s = c % UM_NP;
                    mbar_wait(&s_stage_full[s], (uint32_t)(c / UM_NP) & 1u);
                    tc_fence_after();
                    if (lane == 0) {
                        const int2 win = s_chunk[s];
                        const uint32_t st = smem_u32(stages) + (uint32_t)s * UM_STAGE_BYTES;
                        const uint64_t a_hi = umma_desc(st, 16 * 128, 128), a_lo = umma_desc(st + UM_A_BYTES, 16 * 128, 128);
                        const uint64_t b_hi = umma_desc(st + 2 * UM_A_BYTES, 32 * 128, 128);
                        const uint64_t b_lo = umma_desc(st + 2 * UM_A_BYTES + UM_B_BYTES, 32 * 128, 128);
                        const uint32_t idesc = umma_idesc(win.y);
                        const uint32_t d = tm + (uint32_t)(h * 256 + win.x);
                        umma(d, a_hi, b_hi, idesc, j > 0);
                        umma(d, a_hi, b_lo, idesc, 1);
                        umma(d, a_lo, b_hi, idesc, 1);
                        umma_commit(&s_stage_empty[s]);                       // the stage may be refilled
                        if (j == n_chunks - 1) umma_commit(&s_half_full[h]);  // the half is complete
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
    } else if (warp < 2 + UM_EPI) {
        // =============================== epilogue ================================================================
        const int q = warp & 3;  // tensor-memory lane quarter this warp may read
        float *stg = reinterpret_cast<float *>(epi + (size_t)(warp - 2) * UM_EPI_BYTES);
        const uint32_t stg_s = smem_u32(stg);
        const int n_ct = (W + 31) >> 5;  // 32-column tiles
        for (int k = 0;; ++k) {
            const int slot = k & 1;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k >> 1) & 1u);
            const UmHeader *hd = slot_header(slot);
            const int t = hd->t;
            if (t < 0) break;
            const bool norm = p.normalize && hd->n_live > 0;  // no spot in frame: zeros, returned un-normalised
            float scale = 1.f, vmax = INFINITY;
            if (norm) {
                float m = -INFINITY;
                for (int h = 0; h < n_halves; ++h) {
                    mbar_wait(&s_half_full[h], (uint32_t)k & 1u);
                    tc_fence_after();
                    const int row = 128 * h + 32 * q + lane;
                    for (int ct = 0; ct < n_ct; ++ct) {
                        float v[32];
                        tmem_ld32(tm + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * 256 + 32 * ct), v);
                        if (row < H) {
                            if (32 * ct + 32 <= W) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) m = fmaxf(m, v[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (32 * ct + j < W) m = fmaxf(m, v[j]);
                            }
                        }
                    }
                }
                m = warp_max(m);
                if (lane == 0) s_emax[k & 1][warp - 2] = m;
                asm volatile("bar.sync 1, %0;" ::"n"(UM_EPI * 32) : "memory");
#pragma unroll
                for (int e = 0; e < UM_EPI; ++e) m = fmaxf(m, s_emax[k & 1][e]);
                vmax = m;
                scale = 1.f / vmax;  // np.divide(pattern, np.max(pattern)), simulation2d.py:440-441
            }
            float *img = p.images + (size_t)t * H * W;
            for (int h = 0; h < n_halves; ++h) {
                if (!norm) {
                    mbar_wait(&s_half_full[h], (uint32_t)k & 1u);
                    tc_fence_after();
                }
                const int row0 = 128 * h + 32 * q;  // first image row of this warp's 32 lanes
                for (int ct = 0; ct < n_ct; ++ct) {
                    float v[32];
                    tmem_ld32(tm + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * 256 + 32 * ct), v);
                    if (ct == n_ct - 1) {  // last read of this half by this warp: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s_half_empty[h]);
                    }
                    if (row0 >= H) continue;
                    // numpy divides, so the maximum pixel is exactly 1: pin it (x * (1 / max) can be 1 ulp off)
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (norm && v[j] == vmax) ? 1.0f : v[j] * scale;
                    __syncwarp();  // the previous tile has been read out of the staging buffer
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_s + (uint32_t)(lane * UM_EPI_PITCH + 4 * j) * 4u),
                                     "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                     : "memory");
                    __syncwarp();
                    // 4 rows x 128 contiguous bytes per store instruction
                    const int cg = lane & 7, x = 32 * ct + 4 * cg;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + (lane >> 3), y = row0 + r;
                        const float4 o = lds128v(stg_s + (uint32_t)(r * UM_EPI_PITCH + 4 * cg) * 4u);
                        if (y < H && x < W) {
                            float *d = img + (size_t)y * W + x;
                            if (VEC) {
                                __stcs(reinterpret_cast<float4 *>(d), o);
                            } else {
                                d[0] = o.x;
                                if (x + 1 < W) d[1] = o.y;
                                if (x + 2 < W) d[2] = o.z;
                                if (x + 3 < W) d[3] = o.w;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
    } else {
        // =============================== operand producers =======================================================
        const int pw = warp - (2 + UM_EPI);  // producer index = its stage
        const uint32_t st = smem_u32(stages) + (uint32_t)pw * UM_STAGE_BYTES;
        const int k8 = lane & 7, gq = lane >> 3;
        LutRef L;
        L.base = smem_u32(lut);
        L.n4 = p.n4;
        L.last = p.n4 - 1;
        L.bias = R + LUT_PAD;
        const uint32_t bt_hi = smem_u32(btab), bt_lo = bt_hi + (uint32_t)(8 * n8 * 16);
        int c = 0;
        for (int k = 0;; ++k) {
            const int slot = k & 1;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k >> 1) & 1u);
            const UmHeader *hd = slot_header(slot);
            if (hd->t < 0) break;
            const uint32_t spot_s = smem_u32(slot_spots(slot));
            for (int h = 0; h < n_halves; ++h) {
                const int n_h = hd->n_half[h];
                const int n_chunks = max(1, (n_h + 15) >> 4);
                const uint32_t list_s = smem_u32(slot_list(slot, h));
                for (int j = 0; j < n_chunks; ++j, ++c) {
                    if (c % UM_NP != pw) continue;
                    // this lane's two spots of the chunk (K groups 0 and 1)
                    int sx[2], sy[2];
                    float amp[2];
                    bool valid[2];
                    int xmin = 1 << 20, xmax = -1;
#pragma unroll
                    for (int kh = 0; kh < 2; ++kh) {
                        const int li = 16 * j + 8 * kh + k8;
                        valid[kh] = li < n_h;
                        sx[kh] = sy[kh] = 0;
                        amp[kh] = 0.f;
                        if (valid[kh]) {
                            const uint2 r = lds64v(spot_s + 8u * lds16(list_s + 2u * (uint32_t)li));
                            sx[kh] = (int)(r.x & 0xffffu);
                            sy[kh] = (int)(r.x >> 16);
                            amp[kh] = __uint_as_float(r.y);
                            xmin = min(xmin, sx[kh]);
                            xmax = max(xmax, sx[kh]);
                        }
                    }
                    // column window of the chunk (the first chunk of a half covers every column: it zeroes them)
                    int col0 = 0, ncols = Wp;
                    if (window && j > 0) {
                        xmin = -warp_max(-xmin);
                        xmax = warp_max(xmax);
                        col0 = max(0, xmin - R) & ~15;
                        ncols = ((min(W, xmax + R + 1) - col0) + 15) & ~15;
                    }
                    mbar_wait(&s_stage_empty[pw], ((uint32_t)(c / UM_NP) & 1u) ^ 1u);
#pragma unroll
                    for (int kh = 0; kh < 2; ++kh) {
                        // ---- A: a_s Wy_s[y], rows 128 h + 8 g .. + 7
                        const uint32_t a_hi = st + (uint32_t)(kh * (16 * 128) + k8 * 16), a_lo = a_hi + UM_A_BYTES;
                        const bool fold_y = sy[kh] < R || sy[kh] >= H - R;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int g = gq + 4 * i, y_lo = 128 * h + 8 * g;
                            if (valid[kh] && y_lo + 7 >= sy[kh] - R && y_lo <= sy[kh] + R) {
                                float4 w0, w1;
                                if (fold_y) {
                                    w0 = folded4s(L, y_lo + L.bias, sy[kh], H, R);
                                    w1 = folded4s(L, y_lo + 4 + L.bias, sy[kh], H, R);
                                } else {
                                    const uint32_t a = fetch_addr(L, y_lo + L.bias - sy[kh]);
                                    w0 = lds128(a);
                                    w1 = lds128(a + 16u);
                                }
                                store_split8(a_hi + 128u * g, a_lo + 128u * g, w0, w1, amp[kh]);
                            } else {
                                sts128(a_hi + 128u * g, 0u, 0u, 0u, 0u);
                                sts128(a_lo + 128u * g, 0u, 0u, 0u, 0u);
                            }
                        }
                        // ---- B: Wx_s[x], columns col0 + 8 g .. + 7
                        const uint32_t b_hi = st + 2 * UM_A_BYTES + (uint32_t)(kh * (32 * 128) + k8 * 16), b_lo = b_hi + UM_B_BYTES;
                        const bool fold_x = sx[kh] < R || sx[kh] >= W - R;
                        for (int g = gq; g < (ncols >> 3); g += 4) {
                            const int x_lo = col0 + 8 * g;
                            if (valid[kh] && x_lo + 7 >= sx[kh] - R && x_lo <= sx[kh] + R) {
                                if (fold_x) {
                                    const float4 w0 = folded4s(L, x_lo + L.bias, sx[kh], W, R);
                                    const float4 w1 = folded4s(L, x_lo + 4 + L.bias, sx[kh], W, R);
                                    store_split8(b_hi + 128u * g, b_lo + 128u * g, w0, w1, 1.0f);
                                } else {  // interior: a shifted copy of the pre-split table
                                    const int a = x_lo - sx[kh] + R + UM_BPAD;
                                    const uint32_t off = (uint32_t)(((a & 7) * n8 + (a >> 3)) << 4);
                                    const uint4 vh = lds128u(bt_hi + off), vl = lds128u(bt_lo + off);
                                    sts128(b_hi + 128u * g, vh.x, vh.y, vh.z, vh.w);
                                    sts128(b_lo + 128u * g, vl.x, vl.y, vl.z, vl.w);
                                }
                            } else {
                                sts128(b_hi + 128u * g, 0u, 0u, 0u, 0u);
                                sts128(b_lo + 128u * g, 0u, 0u, 0u, 0u);
                            }
                        }
                    }
                    if (lane == 0) s_chunk[pw] = make_int2(col0, ncols);
                    proxy_fence();  // the stores above become visible to the tensor core's (async proxy) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_stage_full[pw]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

// Returns 1 if the tensor-core kernel was launched, 0 if the configuration is not eligible, < 0 on error.
int launch_render_umma(RenderParams p, cudaStream_t st) {
    if (p.H > 256 || p.W > 256 || p.cap > UM_MAX_CAP || p.radius >= p.W || p.radius >= p.H || p.radius > 120) return 0;
    const int n8 = um_n8(p.radius);
    const int slot_bytes = (32 + p.cap * 8 + 2 * p.cap * 2 + 15) & ~15;
    if (p.cap > 256) p.stage = 0;  // large rows are read in place
    const int front_bytes = (int)(((p.stage ? (size_t)2 * p.cap * 32 : 0) + (size_t)p.table_size * 8 + (size_t)p.cap * 8 +
                                   (size_t)(p.W + 1 + 3) * 4 + 15) & ~(size_t)15);
    const size_t smem = (size_t)UM_NP * UM_STAGE_BYTES + (size_t)UM_EPI * UM_EPI_BYTES + lut_smem_bytes(p.n4) +
                        (size_t)2 * 8 * n8 * 16 + front_bytes + (size_t)2 * slot_bytes;
    if (smem > 226 * 1024) return 0;  // (+ ~1 KB of static shared memory: barriers, alignment)
    const bool vec = (p.W & 3) == 0;
    void (*kern)(RenderParams, int, int, int) = vec ? render_umma_kernel<true> : render_umma_kernel<false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    const int window = option(OPT_RENDER_UMMA_WINDOW) == 0 ? 0 : 1;
    const int sms = num_sms();
    const int grid = p.n_tmpl < sms ? p.n_tmpl : sms;
    kern<<<grid, UM_THREADS, smem, st>>>(p, slot_bytes, front_bytes, window);
    const int rc = check_launch("ds_render (tcgen05)");
    return rc == 0 ? 1 : rc;
}

}  // namespace ds
