// K3 for dense templates: the separable blur as ONE banded matrix product per template on tcgen05.
//
// scipy's gaussian_filter is two 1-D passes.  Here the pass along x is applied analytically while the spots are binned
// by detector row,
//     P[r][x] = sum_{s : row(s) = r} a_s Wx_s[x]            (reflect-folded 1-D weights, render.cu)
// and the pass along y is the constant banded Toeplitz matrix  out[y][x] = sum_r G[y - r] P_ext[r][x],  G[d] = w[|d|],
// where P_ext extends P by mirrored rows (mode="reflect": P_ext[-1 - r] = P[r], P_ext[2H - 1 - r] = P[r]) so that the
// borders need no special operand.  The cost of a template no longer depends on its number of reflections:
//   * K runs over EXTENDED ROWS e = r + RP (RP = radius rounded up to 16), 16 at a time; a chunk of 16 rows is skipped when
//     none of them holds a reflection (sparse templates touch a few chunks only);
//   * B = P_ext chunk (16 x W), written ONCE per chunk to shared memory as bf16 high and low parts (MN-major, no swizzle)
//     by sixteen producer warps, one row each: lane = 8-pixel unit, the row's reflections are summed in float32 in a
//     warp-uniform loop, and the row is computed BEFORE the warp waits for the stage to be free;
//   * A = G[y - r] (128 x 16) never changes: every (half, chunk) pair reads a window of ONE master operand
//     M[yy][k] = G[yy - k] that sits in shared memory for the CTA's lifetime -- the window's first row 128 h + RP - 16 c
//     is a multiple of 16, i.e. a whole number of 8-row core matrices, so it is only an address in the descriptor;
//   * a chunk feeds both halves of the template where its rows reach both (two accumulators of 256 columns in tensor
//     memory), three products per (half, chunk): A_hi B_hi + A_hi B_lo + A_lo B_hi (float32 accumulation);
//   * a half that no chunk reaches is zeroed by one product with an all-zero window of the master.
// 28 warps: front, MMA issue, 16 producers (one per row of a chunk), 8 epilogue; setmaxnreg gives the epilogue warpgroups
// 112 registers and everyone else 56.  Front warp, epilogue (tcgen05.ld -> max -> scale -> swizzled staging -> TMA store),
// slots and barriers are those of render_umma.cu; the per-template records come from render_prepare_rows_kernel below (float64 projection, ordering by
// row with ties in list order, "last write wins" inside a pixel, row offsets, mask of the non-empty chunks).
//
// Reference: diffsims/pattern/detector_functions.py:293-300, diffsims/simulations/simulation2d.py:261-285, :422-441.
#include "umma_device.cuh"

namespace ds {

constexpr int RW_NP_MAX = 6;                      // B stages (chunks in flight between the producers and the tensor core): 3 .. 6
constexpr int RW_EPI = 8;                         // epilogue warps: two per tensor-memory lane quarter
constexpr int RW_PROD = 16;                       // producer warps: one per row of a chunk, all on the same chunk
constexpr int RW_WARPS = 28;                      // seven warpgroups; warps 26, 27 only fill the last one
constexpr int RW_THREADS = RW_WARPS * 32;
constexpr int RW_EPI_WARP0 = 4;                   // warps: 0 front, 1 MMA issue, 4-11 epilogue, 2-3 and 12-25 producers
// B operand in shared memory: 8 x 8 core matrices of 128 contiguous bytes (row k of the K group at 16 k: eight
// consecutive pixels), core matrices of consecutive 8-pixel groups 144 bytes apart -- the 16 bytes of padding spread a
// warp's store of ONE row (lane = pixel group) over all banks; with the natural 128-byte pitch it is a 32-way conflict
constexpr int RW_SBO = 144;
constexpr int RW_LBO = 32 * RW_SBO;               // K groups (rows 0-7 / 8-15 of a chunk)
constexpr int RW_B_BYTES = 2 * RW_LBO;            // one of B_hi / B_lo
constexpr int RW_STAGE_BYTES = 2 * RW_B_BYTES;    // 18 KB
constexpr int RW_TILE_BYTES = 32 * 32 * 4;
constexpr int RW_SLOTS = 4;
constexpr int RW_MAX_CAP = 2048;

#ifdef DS_PROF
// per-role cycle counters of CTA 0 (profiling builds only): [role][0] = cycles in the role's loop, [1 + i] = cycles in wait i
__device__ unsigned long long g_rw_prof[4][4];
#define RPROF_DECL long long rp_t0 = clock64(), rp_w[3] = {0, 0, 0}, rp_s = 0
#define RPROF_BEGIN rp_s = clock64()
#define RPROF_END(i) rp_w[i] += clock64() - rp_s
#define RPROF_DONE(role)                                                                 \
    if (blockIdx.x == 0 && lane == 0) {                                                  \
        g_rw_prof[role][0] = (unsigned long long)(clock64() - rp_t0);                    \
        for (int i_ = 0; i_ < 3; ++i_) g_rw_prof[role][1 + i_] = (unsigned long long)rp_w[i_]; \
    }
#else
#define RPROF_DECL
#define RPROF_BEGIN
#define RPROF_END(i)
#define RPROF_DONE(role)
#endif

struct RowsHeader {  // 32 bytes at the start of a record / slot
    int t, n_live;
    unsigned mask;  // bit c: extended rows 16 c .. 16 c + 15 hold at least one reflection
    int pad[5];
};

// geometry of the extended rows and of the master operand for kernel radius R
struct RowsGeom {
    int RP;           // radius rounded up to a multiple of 16: extended row e = r + RP
    int d_min;        // first master row (master row yy holds G[yy - k], k = 0..15)
    int d_zero;       // a window starting here is all zero
    int master_rows;  // multiple of 16
    int n_chunks;     // extended-row chunks of a template
};
__host__ __device__ inline RowsGeom rows_geom(int R, int H) {
    RowsGeom g;
    g.RP = (R + 15) & ~15;
    g.d_min = g.RP - 16 * ((127 + R + g.RP) / 16);
    g.d_zero = g.RP + 16;
    g.master_rows = 128 + g.d_zero - g.d_min;
    g.n_chunks = (((H + 15) & ~15) + 2 * g.RP) / 16;
    return g;
}
// source row of extended row e (-1: the row lies beyond the reach of the kernel or of the image)
__device__ __forceinline__ int rows_source(int e, int RP, int R, int H) {
    const int r = e - RP;
    if (r < -R || r > H - 1 + R) return -1;
    return r < 0 ? -1 - r : (r >= H ? 2 * H - 1 - r : r);
}

// ---------------------------------------------------------------------------------------------------
// per-template preparation, one warp per template
// ---------------------------------------------------------------------------------------------------
constexpr int PREPR_BINS = 9 * 32;  // row histogram entries (H <= 256 rows + 1, padded to 9 per lane)

__global__ void __launch_bounds__(256) render_prepare_rows_kernel(const RenderParams p, unsigned char *records, const int record_bytes,
                                                                  const int warp_bytes) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    unsigned char *base = smem_raw + (size_t)warp * warp_bytes;
    // (shared memory per warp bounds the warps per SM of this latency-bound pass: the amplitudes are re-read from global
    // memory at scatter time instead of being kept here)
    uint2 *sspot = reinterpret_cast<uint2 *>(base);                                  // [cap] sorted spots
    unsigned *pkey = reinterpret_cast<unsigned *>(base + (size_t)p.cap * 8);         // [cap] column | row << 16 of spot j, ~0 = not in frame
    int *bins = reinterpret_cast<int *>(base + (size_t)p.cap * 12);
    const int R = p.radius, H = p.H, W = p.W;
    const RowsGeom geo = rows_geom(R, H);

    for (int t = blockIdx.x * n_warps + warp; t < p.n_tmpl; t += gridDim.x * n_warps) {
        const int n = min(p.count[t], p.cap);
        const double *sxyz = p.xyz + (size_t)t * p.cap * 3;
        const double *sint = p.intensity + (size_t)t * p.cap;
        unsigned char *rec = records + (size_t)t * record_bytes;
        RowsHeader *hd = reinterpret_cast<RowsHeader *>(rec);
        uint2 *gspot = reinterpret_cast<uint2 *>(rec + 32);
        unsigned short *goff = reinterpret_cast<unsigned short *>(rec + rows_offsets_offset(p.cap));
        // ---- histogram of the in-frame spots over detector rows
        for (int e = lane; e < PREPR_BINS; e += 32) bins[e] = 0;
        __syncwarp();
        int n_live = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            unsigned kk = 0xffffffffu;
            if (j < n) {
                double px, py;
                project_spot(p, sxyz[3 * j], sxyz[3 * j + 1], px, py);
                if (px >= 0.0 && px < (double)W && py >= 0.0 && py < (double)H)  // astype(int) truncates
                    kk = (unsigned)(int)px | ((unsigned)(int)py << 16);
                if (kk != 0xffffffffu) atomicAdd(&bins[(kk >> 16) + 1], 1);
                pkey[j] = kk;
            }
            n_live += __popc(__ballot_sync(0xffffffffu, kk != 0xffffffffu));
        }
        __syncwarp();
        {   // exclusive scan of bins[0 .. H]: each lane owns 9 consecutive entries (H + 1 <= 288)
            int loc[9], sum = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int e = 9 * lane + i;
                loc[i] = e <= H ? bins[e] : 0;
                sum += loc[i];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            int run = incl - sum;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int e = 9 * lane + i;
                run += loc[i];
                if (e <= H) bins[e] = run;  // inclusive over bins[0 .. e] = first slot of row e
            }
        }
        __syncwarp();
        // ---- stable scatter: 32 spots at a time, the lanes of one row ranked by lane number; bins[r] ends as the end of row r
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const unsigned kk = j < n ? pkey[j] : 0xffffffffu;
            const unsigned mask = __ballot_sync(0xffffffffu, kk != 0xffffffffu);
            int row = 0, at = 0;
            unsigned peers = 0;
            if (kk != 0xffffffffu) {
                row = (int)(kk >> 16);
                peers = __match_any_sync(mask, row);
                at = bins[row] + __popc(peers & ((1u << lane) - 1u));
            }
            __syncwarp();
            if (kk != 0xffffffffu) {
                sspot[at] = make_uint2(kk, __float_as_uint((float)sint[j]));
                if ((peers >> lane) == 1u) bins[row] = at + 1;  // the highest lane of the row leaves its end
            }
            __syncwarp();
        }
        // ---- last write wins: a spot is overwritten if a later spot of its row (they follow it directly, in list order)
        // sits in the same column
        for (int i0 = 0; i0 < n_live; i0 += 32) {
            const int i = i0 + lane;
            if (i < n_live) {
                uint2 v = sspot[i];
                for (int i2 = i + 1; i2 < n_live; ++i2) {
                    const unsigned k2 = sspot[i2].x;
                    if ((k2 >> 16) != (v.x >> 16)) break;
                    if (k2 == v.x) {
                        v.y = 0u;
                        break;
                    }
                }
                gspot[i] = v;
            }
        }
        // ---- row offsets (row r holds spots off[r] .. off[r + 1] - 1) and the mask of non-empty extended chunks
        for (int r = lane; r <= H; r += 32) goff[r] = (unsigned short)(r == 0 ? 0 : bins[r - 1]);
        bool any = false;
        if (lane < geo.n_chunks)
            for (int i = 0; i < 16; ++i) {
                const int src = rows_source(16 * lane + i, geo.RP, R, H);
                if (src >= 0) any |= bins[src] > (src == 0 ? 0 : bins[src - 1]);
            }
        const unsigned cmask = __ballot_sync(0xffffffffu, any);
        if (lane == 0) {
            hd->t = t;
            hd->n_live = n_live;
            hd->mask = cmask;
        }
        __syncwarp();
    }
}

static int launch_render_prepare_rows(const RenderParams &p, unsigned char *records, cudaStream_t st) {
    const int record_bytes = rows_record_bytes(p.cap);
    const int warp_bytes = (p.cap * 12 + PREPR_BINS * 4 + 15) & ~15;
    int warps = 8;
    while (warps > 1 && (size_t)warps * warp_bytes > 110 * 1024) warps >>= 1;   // (two CTAs per SM)
    const size_t smem = (size_t)warps * warp_bytes;
    if (smem > 48 * 1024) cudaFuncSetAttribute(render_prepare_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    const int want = (p.n_tmpl + warps - 1) / warps;
    const int grid = want < 16 * num_sms() ? want : 16 * num_sms();
    render_prepare_rows_kernel<<<grid, warps * 32, smem, st>>>(p, records, record_bytes, warp_bytes);
    return check_launch("ds_render (prepare rows)");
}

// ---------------------------------------------------------------------------------------------------
// the render kernel
// ---------------------------------------------------------------------------------------------------
template <int RW_NP>
__global__ void __launch_bounds__(RW_THREADS, 1) render_rows_kernel(const RenderParams p, const __grid_constant__ CUtensorMap tmap,
                                                                     const unsigned char *records, const int slot_bytes,
                                                                     const int epi_bufs) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_slot_full[RW_SLOTS], s_slot_empty[RW_SLOTS], s_stage_full[RW_NP_MAX], s_stage_empty[RW_NP_MAX],
        s_half_full[2], s_half_empty[2];
    __shared__ float s_emax[2][RW_EPI];
    __shared__ uint32_t s_tmem;
    __shared__ double s_norm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.radius, H = p.H, W = p.W;
    const int n_halves = (H + 127) >> 7;
    const int Wp = (W + 15) & ~15;  // columns of a product
    const RowsGeom geo = rows_geom(R, H);
    const int MG = geo.master_rows >> 3;         // 8-row groups of the master operand
    const uint32_t master_part = (uint32_t)(2 * MG * 128);  // bytes of master hi (and of master lo)

    // ---- shared memory: B stages | epilogue tiles (1024-byte aligned) | master hi | master lo | tap LUT | slots
    unsigned char *stages = smem_raw;
    unsigned char *epi = stages + (size_t)RW_NP * RW_STAGE_BYTES;
    unsigned char *master = epi + (size_t)RW_EPI * epi_bufs * RW_TILE_BYTES;
    float4 *lut = reinterpret_cast<float4 *>(master + 2 * (size_t)master_part);
    unsigned char *slots = reinterpret_cast<unsigned char *>(lut) + lut_smem_bytes(p.n4);
    auto slot_header = [&](int s) { return reinterpret_cast<RowsHeader *>(slots + (size_t)s * slot_bytes); };

    // chunk ranges of the halves: extended rows 128 h - R + RP .. min(H - 1, 128 h + 127) + R + RP
    unsigned range[2] = {0u, 0u};
    for (int h = 0; h < n_halves; ++h) {
        const int lo = (128 * h - R + geo.RP) >> 4, hi = min(geo.n_chunks - 1, (min(H - 1, 128 * h + 127) + R + geo.RP) >> 4);
        range[h] = (hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
    }

    // ---- CTA-wide, once ---------------------------------------------------------------------------------------
    if (warp == 0) {
        double part = 0.0;
        for (int k = lane; k <= R; k += 32) part += (k == 0 ? 1.0 : 2.0) * exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
            s_norm = part;
            for (int s = 0; s < RW_SLOTS; ++s) {
                mbar_init(&s_slot_full[s], 1);
                mbar_init(&s_slot_empty[s], 1 + RW_EPI + RW_PROD);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&s_half_full[s], 1);
                mbar_init(&s_half_empty[s], RW_EPI);
            }
            for (int s = 0; s < RW_NP; ++s) {
                mbar_init(&s_stage_full[s], RW_PROD);
                mbar_init(&s_stage_empty[s], 1);
            }
            fence_mbar_init();
        }
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp == RW_EPI_WARP0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    fill_lut(lut, p.n4, R, p.sigma, 1.0 / s_norm, threadIdx.x, RW_THREADS);
    {   // master operand M[yy][k] = G[yy - k], yy = d_min .. d_min + master_rows - 1, k = 0 .. 15, as bf16 hi and lo in the
        // canonical MN-major layout: K group k >> 3, 8-row group, then 16 bytes per k holding 8 consecutive rows
        const double inv_norm = 1.0 / s_norm;
        for (int e = threadIdx.x; e < geo.master_rows * 16; e += RW_THREADS) {
            const int k = e & 15, yi = e >> 4;
            const int dlt = abs(yi + geo.d_min - k);
            const float w = (dlt <= R) ? (float)(exp(-0.5 / (p.sigma * p.sigma) * (double)dlt * (double)dlt) * inv_norm) : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            const size_t off = (size_t)(k >> 3) * MG * 128 + (size_t)(yi >> 3) * 128 + (size_t)(k & 7) * 16 + (size_t)(yi & 7) * 2;
            *reinterpret_cast<unsigned short *>(master + off) = __bfloat16_as_ushort(hi);
            *reinterpret_cast<unsigned short *>(master + master_part + off) = __bfloat16_as_ushort(lo);
        }
    }
    proxy_fence();  // the master is read by the tensor core (async proxy)
    __syncthreads();
    const uint32_t tm = s_tmem;

    // Role dispatch.  The kernel starts with 72 registers per thread (896 threads); the two epilogue warpgroups (warps
    // 4..11) raise their allowance to 112 and every other warpgroup lowers it to 56 (an increase can only draw on what the
    // CTA's own warps released) -- each setmaxnreg sits at the head of the branch it governs (one instruction per
    // warpgroup, and ptxas allocates each branch against its own limit).
    const bool epi_role = warp >= RW_EPI_WARP0 && warp < RW_EPI_WARP0 + RW_EPI;
    if (!epi_role) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // =============================== front warp ==============================================================
        RPROF_DECL;
        int t_ahead = 0;
        if (lane == 0) t_ahead = atomicAdd(&p.ticket[0], 1);
        for (int k = 0;; ++k) {
            const int slot = k % RW_SLOTS;
            const int t = __shfl_sync(0xffffffffu, t_ahead, 0);
            if (lane == 0 && t < p.n_tmpl) t_ahead = atomicAdd(&p.ticket[0], 1);
            RPROF_BEGIN;
            mbar_wait(&s_slot_empty[slot], ((uint32_t)(k / RW_SLOTS) & 1u) ^ 1u);
            RPROF_END(0);
            if (t >= p.n_tmpl) {  // out of work: stop slot
                if (lane == 0) {
                    slot_header(slot)->t = -1;
                    mbar_arrive(&s_slot_full[slot]);
                }
                RPROF_DONE(0);
                break;
            }
            if (lane == 0) {  // the copy's completion is the slot's `full` signal
                mbar_expect_tx(&s_slot_full[slot], (uint32_t)slot_bytes);
                bulk_g2s(slots + (size_t)slot * slot_bytes, records + (size_t)t * slot_bytes, (uint32_t)slot_bytes, &s_slot_full[slot]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =============================== MMA issue ===============================================================
        RPROF_DECL;
        int stage = 0;
        uint32_t sphase = 0;
        const uint32_t st0 = smem_u32(stages), ms = smem_u32(master);
        const uint32_t lbo_a = (uint32_t)MG * 128u;
        const uint64_t d_b_hi = umma_desc(st0, RW_LBO, RW_SBO), d_b_lo = umma_desc(st0 + RW_B_BYTES, RW_LBO, RW_SBO);
        const uint32_t idesc = umma_idesc(Wp);
        const uint64_t d_a_hi0 = umma_desc(ms, lbo_a, 128), d_a_lo0 = umma_desc(ms + master_part, lbo_a, 128);
        for (int k = 0;; ++k) {
            const int slot = k % RW_SLOTS;
            RPROF_BEGIN;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k / RW_SLOTS) & 1u);
            RPROF_END(0);
            const RowsHeader *hd = slot_header(slot);
            if (hd->t < 0) {
                RPROF_DONE(1);
                break;
            }
            const unsigned mask = hd->mask;
            unsigned m[2] = {mask & range[0], mask & range[1]};
            if (elect_one()) {
                // a half no chunk reaches: one product with the all-zero window of the master (B: any finite operand)
                for (int h = 0; h < n_halves; ++h)
                    if (m[h] == 0u) {
                        mbar_wait(&s_half_empty[h], ((uint32_t)k & 1u) ^ 1u);
                        tc_fence_after();
                        umma(tm + (uint32_t)(h * 256), umma_desc(ms + (uint32_t)((geo.d_zero - geo.d_min) >> 3) * 128u, lbo_a, 128),
                             umma_desc(ms, lbo_a, 128), idesc, 0);
                        umma_commit(&s_half_full[h]);
                    }
                unsigned all = m[0] | m[1];
                while (all) {
                    const int c = __ffs(all) - 1;
                    all &= all - 1;
                    RPROF_BEGIN;
                    mbar_wait(&s_stage_full[stage], sphase);
                    RPROF_END(1);
                    tc_fence_after();
                    const uint64_t so = (uint64_t)((uint32_t)stage * (RW_STAGE_BYTES >> 4));
                    for (int h = 0; h < n_halves; ++h) {
                        if (!((m[h] >> c) & 1u)) continue;
                        const bool first = (m[h] & ((1u << c) - 1u)) == 0u, last = (m[h] >> c) == 1u;
                        if (first) {  // the epilogue has drained template k - 1
                            RPROF_BEGIN;
                            mbar_wait(&s_half_empty[h], ((uint32_t)k & 1u) ^ 1u);
                            RPROF_END(2);
                            tc_fence_after();
                        }
                        // window of the master: rows 128 h + RP - 16 c .. + 127 = 8-row group 16 h - 2 c + (RP - d_min) / 8
                        // (128 bytes each: 8 descriptor address units)
                        const uint64_t wo = (uint64_t)(uint32_t)(8 * (16 * h - 2 * c + ((geo.RP - geo.d_min) >> 3)));
                        const uint64_t d_a_hi = d_a_hi0 + wo, d_a_lo = d_a_lo0 + wo;
                        const uint32_t d = tm + (uint32_t)(h * 256);
                        umma_keep_a(d, d_a_hi, d_b_hi + so, idesc, first ? 0u : 1u);
                        umma_reuse_a(d, d_a_hi, d_b_lo + so, idesc, 1);
                        umma(d, d_a_lo, d_b_hi + so, idesc, 1);
                        if (last) umma_commit(&s_half_full[h]);  // the half is complete
                    }
                    umma_commit(&s_stage_empty[stage]);  // the stage may be refilled
                    if (++stage == RW_NP) {
                        stage = 0;
                        sphase ^= 1u;
                    }
                }
            }
            // (the other lanes only follow the template structure)
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
    } else if ((warp < RW_EPI_WARP0 ? warp - 2 : warp - (RW_EPI_WARP0 + RW_EPI) + 2) < RW_PROD) {
        // =============================== B producers ==============================================================
        // All sixteen warps work on the same chunk, one row each (the time to fill a stage is one row's latency); lane =
        // 8-pixel unit, so the row's reflections are a warp-uniform loop and a reflection costs two table loads and
        // eight FMAs per lane.
        const int pi = warp < RW_EPI_WARP0 ? warp - 2 : warp - (RW_EPI_WARP0 + RW_EPI) + 2;  // 0 .. RW_PROD - 1 = row of the chunk
        LutRef L;
        L.base = smem_u32(lut);
        L.n4 = p.n4;
        L.last = p.n4 - 1;
        L.bias = R + LUT_PAD;
        const int x_lo = 8 * lane;
        const bool unit_ok = x_lo < Wp;
        // row pi of the chunk: K group pi >> 3, 16 (pi & 7) inside the core matrix
        const uint32_t b_row = smem_u32(stages) + (uint32_t)((pi >> 3) * RW_LBO + (pi & 7) * 16) + (uint32_t)lane * RW_SBO;
        RPROF_DECL;
        int c = 0;
        for (int k = 0;; ++k) {
            const int slot = k % RW_SLOTS;
            RPROF_BEGIN;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k / RW_SLOTS) & 1u);
            RPROF_END(0);
            const RowsHeader *hd = slot_header(slot);
            if (hd->t < 0) {
                if (pi == 0) { RPROF_DONE(3); }
                break;
            }
            const uint32_t spot_s = smem_u32(slots + (size_t)slot * slot_bytes + 32);
            const uint32_t off_s = smem_u32(slots + (size_t)slot * slot_bytes + rows_offsets_offset(p.cap));
            unsigned all = hd->mask & (range[0] | range[1]);
            for (; all; all &= all - 1, ++c) {
                const int ce = __ffs(all) - 1;
                const int stage = c % RW_NP;
                const uint32_t b_hi = b_row + (uint32_t)stage * RW_STAGE_BYTES, b_lo = b_hi + RW_B_BYTES;
                // this warp's row and its reflections (read before the wait for the stage)
                const int src = rows_source(16 * ce + pi, geo.RP, R, H);
                int s0 = 0, s1 = 0;
                if (src >= 0) {
                    s0 = (int)lds16(off_s + 2u * (uint32_t)src);
                    s1 = (int)lds16(off_s + 2u * (uint32_t)src + 2u);
                }
                float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
                // += am * taps a .. a + 7 of the padded kernel for this lane's unit, a = a0 + 8 lane: a0 is the same for every
                // lane, so the table copy (a & 3) and the entry ((a >> 2) = (a0 >> 2) + 2 lane) split into a warp-uniform
                // part and 32 lane; lanes outside [lane_lo, lane_hi] read the all-zero head of the table
                auto add_taps = [&](int a0, int lane_lo, int lane_hi, float am) {
                    const uint32_t ad = (lane >= lane_lo && lane <= lane_hi)
                                            ? L.base + (uint32_t)((((a0 & 3) * L.n4 + (a0 >> 2)) << 4) + 32 * lane)
                                            : L.base;
                    const float4 u0 = lds128(ad), u1 = lds128(ad + 16u);
                    w0 = make_float4(fmaf(am, u0.x, w0.x), fmaf(am, u0.y, w0.y), fmaf(am, u0.z, w0.z), fmaf(am, u0.w, w0.w));
                    w1 = make_float4(fmaf(am, u1.x, w1.x), fmaf(am, u1.y, w1.y), fmaf(am, u1.z, w1.z), fmaf(am, u1.w, w1.w));
                };
                uint2 r_next = make_uint2(0u, 0u);
                if (s0 < s1) r_next = lds64v(spot_s + 8u * (uint32_t)s0);  // (the same word for every lane)
                for (int s = s0; s < s1; ++s) {
                    const uint2 r = r_next;
                    if (s + 1 < s1) r_next = lds64v(spot_s + 8u * (uint32_t)(s + 1));
                    const int cx = (int)(r.x & 0xffffu);
                    const float am = __uint_as_float(r.y);
                    // direct image: offsets d = 8 lane - cx in [-R - 7, R]
                    add_taps(L.bias - cx, (cx - R) >> 3, (cx + R) >> 3, am);
                    if (cx < R)  // mirror image at -cx - 1 (warp-uniform): d2 = 8 lane + cx + 1 <= R
                        add_taps(L.bias + cx + 1, 0, (R - cx - 1) >> 3, am);
                    if (cx >= W - R)  // mirror image at 2 W - 1 - cx: d3 = 8 lane + cx + 1 - 2 W in [-R - 7, R]
                        add_taps(L.bias + cx + 1 - 2 * W, (2 * W - cx - 1 - R) >> 3, (2 * W - cx - 1 + R) >> 3, am);
                }
                uint32_t h[4] = {0u, 0u, 0u, 0u}, l[4] = {0u, 0u, 0u, 0u};
                if (s0 != s1) {  // (warp-uniform)
                    bf16_split2(w0.x, w0.y, h[0], l[0]);
                    bf16_split2(w0.z, w0.w, h[1], l[1]);
                    bf16_split2(w1.x, w1.y, h[2], l[2]);
                    bf16_split2(w1.z, w1.w, h[3], l[3]);
                }
                RPROF_BEGIN;
                mbar_wait(&s_stage_empty[stage], ((uint32_t)(c / RW_NP) & 1u) ^ 1u);
                RPROF_END(1);
                if (unit_ok) {
                    sts128(b_hi, h[0], h[1], h[2], h[3]);
                    sts128(b_lo, l[0], l[1], l[2], l[3]);
                }
                proxy_fence();  // the stores above become visible to the tensor core's (async proxy) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_stage_full[stage]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // =============================== epilogue (as in render_umma.cu) ========================================
        const int ew = warp - RW_EPI_WARP0;  // 0..7
        const int q = warp & 3;              // tensor-memory lane quarter this warp may read
        const int ch = ew >> 2;              // of the 32-column tiles of a half this warp takes ct = ch, ch + 2, ...
        const uint32_t tile_s = smem_u32(epi + (size_t)ew * epi_bufs * RW_TILE_BYTES);
        const int n_ct = (W + 31) >> 5;
        const uint32_t tm_q = tm + ((uint32_t)(32 * q) << 16);
        int nbuf = 0;  // staged tiles so far (buffer = nbuf & 1)
        RPROF_DECL;
        for (int k = 0;; ++k) {
            const int slot = k % RW_SLOTS;
            RPROF_BEGIN;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k / RW_SLOTS) & 1u);
            RPROF_END(0);
            const RowsHeader *hd = slot_header(slot);
            const int t = hd->t;
            if (t < 0) {
                if (ew == 0) { RPROF_DONE(2); }
                break;
            }
            const bool norm = p.normalize && hd->n_live > 0;  // no spot in frame: zeros, returned un-normalised
            float scale = 1.f, vmax = INFINITY, clamp = __uint_as_float(0x7fc00000u);
            if (norm) {
                // ---- pass 1: the template maximum, straight from the accumulators
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
                for (int h = 0; h < n_halves; ++h) {
                    RPROF_BEGIN;
                    mbar_wait(&s_half_full[h], (uint32_t)k & 1u);
                    RPROF_END(1);
                    tc_fence_after();
                    if (128 * h + 32 * q >= H) continue;  // none of this warp's rows is in the image (warp-uniform)
                    const bool row_ok = 128 * h + 32 * q + lane < H;
                    uint32_t r[32];
                    if (ch < n_ct) tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * ch), r);
#pragma unroll 1
                    for (int ct = ch; ct < n_ct; ct += 2) {
                        tmem_wait(r);
                        if (row_ok) {
                            if (32 * ct + 32 <= W) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    m0 = fmaxf(m0, __uint_as_float(r[j]));
                                    m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
                                    m2 = fmaxf(m2, __uint_as_float(r[j + 2]));
                                    m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
                                }
                            } else {
#pragma unroll  // (fully: a dynamic index would put the array in local memory)
                                for (int j = 0; j < 32; ++j)
                                    if (32 * ct + j < W) m0 = fmaxf(m0, __uint_as_float(r[j]));
                            }
                        }
                        if (ct + 2 < n_ct) tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * (ct + 2)), r);
                    }
                }
                float m = warp_max(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                if (lane == 0) s_emax[k & 1][ew] = m;
                RPROF_BEGIN;
                asm volatile("bar.sync 1, %0;" ::"n"(RW_EPI * 32) : "memory");
                RPROF_END(2);
#pragma unroll
                for (int e = 0; e < RW_EPI; ++e) m = fmaxf(m, s_emax[k & 1][e]);
                vmax = m;
                // np.divide(pattern, np.max(pattern)), simulation2d.py:440-441: the maximum pixel is exactly 1 there.
                // x * (1 / max) can be 1 ulp off, so the reciprocal is rounded UP and the product clamped to 1; a NaN clamp
                // leaves everything as computed when the maximum is not a positive finite number.
                scale = 1.f / vmax;
                if (vmax > 0.f && vmax < INFINITY) {
                    if (scale * vmax < 1.f) scale = __uint_as_float(__float_as_uint(scale) + 1u);
                    clamp = 1.f;
                }
            }
            // ---- pass 2: scale, stage, store
            auto stage_tile = [&](const uint32_t (&r)[32], int ct, int row0) {
                if (elect_one()) {  // the copy that last read this buffer has finished reading it
                    if (epi_bufs == 2)
                        tma_store_wait_read<1>();
                    else
                        tma_store_wait_read<0>();
                }
                __syncwarp();
                const uint32_t buf = tile_s + (uint32_t)(epi_bufs == 2 ? (nbuf & 1) : 0) * RW_TILE_BYTES;
                // 128B swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4))),
                                 "f"(fminf(__uint_as_float(r[4 * j]) * scale, clamp)), "f"(fminf(__uint_as_float(r[4 * j + 1]) * scale, clamp)),
                                 "f"(fminf(__uint_as_float(r[4 * j + 2]) * scale, clamp)), "f"(fminf(__uint_as_float(r[4 * j + 3]) * scale, clamp))
                                 : "memory");
                proxy_fence();
                __syncwarp();
                if (elect_one()) tma_store_tile(&tmap, buf, 32 * ct, row0, t);
                ++nbuf;
            };
#pragma unroll 1
            for (int h = 0; h < n_halves; ++h) {
                if (!norm) {
                    mbar_wait(&s_half_full[h], (uint32_t)k & 1u);
                    tc_fence_after();
                }
                const int row0 = 128 * h + 32 * q;  // first image row of this warp's 32 lanes
                if (row0 < H && ch < n_ct) {
                    uint32_t ra[32], rb[32];
                    tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * ch), ra);
#pragma unroll 1
                    for (int ct = ch; ct < n_ct; ct += 4) {
                        if (ct + 2 < n_ct) tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * (ct + 2)), rb);
                        tmem_wait(ra);
                        stage_tile(ra, ct, row0);
                        if (ct + 2 < n_ct) {
                            if (ct + 4 < n_ct) tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * (ct + 4)), ra);
                            tmem_wait(rb);
                            stage_tile(rb, ct + 2, row0);
                        }
                    }
                }
                // this warp's reads of the half are complete: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_half_empty[h]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
        if (elect_one()) tma_store_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

bool rows_eligible(int H, int W, int cap, int radius) {
    return !(H > 256 || W > 256 || (W & 3) != 0 || cap > RW_MAX_CAP || radius >= W || radius >= H || radius > 120 || radius < 1);
}

// Returns 1 if the kernel was launched, 0 if the configuration is not eligible, < 0 on error.
// `records`: n_tmpl * rows_record_bytes(cap) bytes of device scratch (the prepared templates).
int launch_render_rows(RenderParams p, unsigned char *records, cudaStream_t st) {
    if (!rows_eligible(p.H, p.W, p.cap, p.radius)) return 0;
    const RowsGeom geo = rows_geom(p.radius, p.H);
    if (geo.n_chunks > 32) return 0;
    const int slot_bytes = rows_record_bytes(p.cap);
    int epi_bufs = 2, n_stages = RW_NP_MAX;
    auto smem_for = [&](int bufs, int stages) {
        return (size_t)stages * RW_STAGE_BYTES + (size_t)RW_EPI * bufs * RW_TILE_BYTES + (size_t)4 * (geo.master_rows >> 3) * 128 +
               lut_smem_bytes(p.n4) + (size_t)RW_SLOTS * slot_bytes;
    };
    // as many stages as fit next to two staged tiles per epilogue warp, at least three
    while (n_stages > 3 && smem_for(2, n_stages) > 226 * 1024) --n_stages;
    {
        const int o = option(OPT_RENDER_ROWS_STAGES);
        if (o >= 3 && o <= RW_NP_MAX && smem_for(2, o) <= 226 * 1024) n_stages = o;
    }
    if (smem_for(2, n_stages) > 226 * 1024) epi_bufs = 1;
    const size_t smem = smem_for(epi_bufs, n_stages);
    if (smem > 226 * 1024) return 0;  // (+ ~1 KB of static shared memory: barriers, alignment)
    alignas(64) CUtensorMap tmap;
    const int rt = make_image_tensor_map(&tmap, p.images, p.n_tmpl, p.H, p.W, 32, 32, true);
    if (rt != 0) return rt < 0 ? rt : 0;
    const int rc0 = launch_render_prepare_rows(p, records, st);
    if (rc0 != 0) return rc0;
    const int sms = num_sms();
    const int grid = p.n_tmpl < sms ? p.n_tmpl : sms;
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        kern<<<grid, RW_THREADS, smem, st>>>(p, tmap, records, slot_bytes, epi_bufs);
    };
    switch (n_stages) {  // (a compile-time stage count: the stage arithmetic sits on every role's critical path)
        case 3: launch(render_rows_kernel<3>); break;
        case 4: launch(render_rows_kernel<4>); break;
        case 5: launch(render_rows_kernel<5>); break;
        default: launch(render_rows_kernel<6>); break;
    }
    const int rc = check_launch("ds_render (tcgen05, rows)");
    return rc == 0 ? 1 : rc;
}

}  // namespace ds

#ifdef DS_PROF
extern "C" int ds_debug_rows_prof(unsigned long long *out_host /*[16]*/) {
    return cudaMemcpyFromSymbol(out_host, ds::g_rw_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif
