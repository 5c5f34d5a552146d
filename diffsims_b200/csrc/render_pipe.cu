// K3, pipelined variant: row capacities up to 512 (the BASELINE headline case and every library whose slots fit
// in shared memory).
//
// Same arithmetic as render_kernel<FAST> (render.cu); different schedule.  In render_kernel all warps of a
// group walk through the phases of a template together (load spots -> project / dedupe -> region bounds ->
// max pass -> store pass), so for a sparse template the store stream of an SM pauses during every
// non-store phase (measured: 0.90 of the HBM peak normalised vs 0.976 without the max pass).  Here the
// CTA is warp-specialised:
//   front warp(s)       prepare template k+1: prefetched spot rows (cp.async.bulk) -> projection, last-write-wins
//                       hash, compaction -> region upper bounds -> max pass over the few regions that can hold
//                       the maximum -> scale; publishes a slot and arrives on its `full` mbarrier;
//   render warps        drain slot k: regions are handed out by a shared-memory ticket, each region is
//                       accumulated in registers and streamed out with st.global.cs.v4; every lane arrives on
//                       the slot's `empty` mbarrier when the ticket runs out.
// Two slots per front warp, so stores of template k overlap the whole preparation of template k+1.
// DENSE instantiations carry the tensor-core path of accumulate_region (render_device.cuh) for libraries with
// hundreds of reflections per template; sparse libraries run the lean ones.
#include <cuda.h>  // CUtensorMap (type only)

#include "render_device.cuh"

namespace ds {

int make_image_tensor_map(CUtensorMap *tmap, float *images, int n_tmpl, int H, int W, int box_w, int box_h, bool swizzle128);

// NF front warps (1 for the sparsest patterns, 2 when the preparation of a template is heavier) feed
// RN_WARPS - NF render warps through 2 NF slots; front f owns the sequence numbers k = f (mod NF).

struct PipeHeader {  // 32 bytes at the start of a slot
    int n_live, n_pass, t, pad0;
    float scale, vmax, pad1, pad2;
};

template <bool VEC, int NF, bool DENSE>
__global__ void __launch_bounds__(RN_THREADS, 2) render_pipe_kernel(const RenderParams p, const int slot_bytes,
                                                                    const int front_bytes,
                                                                    const __grid_constant__ CUtensorMap tmap, const int zero_tma,
                                                                    const int zero_tma_offset) {
    constexpr int RP_SLOTS = 2 * NF;
    constexpr int RP_RENDER_WARPS = RN_WARPS - NF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[RP_SLOTS], s_empty[RP_SLOTS], s_stage_all[NF][2];
    __shared__ int s_ticket[RP_SLOTS];
    __shared__ double s_norm;
    // All-zero regions (most of a sparse template) are not stored by the lanes: one elected lane hands this zero tile
    // to the TMA engine (cp.async.bulk.tensor store, one 64 x 32 box), which keeps the LSU / register path for the
    // regions that carry data.
    // (the tile sits at the END of the dynamic shared memory, only when the option is on)
    float *s_zero = reinterpret_cast<float *>(smem_raw + zero_tma_offset);
    if (zero_tma) {
        for (int e = threadIdx.x; e < RN_RW * RN_RH; e += RN_THREADS) s_zero[e] = 0.f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrx = (p.W + RN_RW - 1) / RN_RW, nry = (p.H + RN_RH - 1) / RN_RH;
    const int n_regions = nrx * nry;

    // ---- shared memory: LUT | front workspace (stage x2, hash, key, inten) | slots ------------------------
    float4 *lut = reinterpret_cast<float4 *>(smem_raw);
    const int front = warp < NF ? warp : 0;
    const size_t shared_prefix = lut_smem_bytes(p.n4) + (size_t)p.hits_bytes;  // LUT | hit lists
    const uint32_t hits_s =
        p.hits_bytes ? smem_u32(smem_raw + lut_smem_bytes(p.n4)) + (uint32_t)(warp * (p.hits_bytes / RN_WARPS)) : 0u;
    unsigned char *base = smem_raw + shared_prefix + (size_t)front * front_bytes;
    uint64_t *s_stage = s_stage_all[front];
    double *stage[2] = {nullptr, nullptr};
    if (p.stage) {
        stage[0] = reinterpret_cast<double *>(base);
        stage[1] = stage[0] + p.cap * 4;
        base += (size_t)2 * p.cap * 32;
    }
    unsigned long long *hash = reinterpret_cast<unsigned long long *>(base);
    base += (size_t)p.table_size * 8;
    int *key = reinterpret_cast<int *>(base);
    base += (size_t)p.cap * 4;
    float *inten = reinterpret_cast<float *>(base);
    base = smem_raw + shared_prefix + (size_t)NF * front_bytes;
    unsigned char *slots = base;
    auto slot_header = [&](int s) { return reinterpret_cast<PipeHeader *>(slots + (size_t)s * slot_bytes); };
    auto slot_spots = [&](int s) { return reinterpret_cast<uint2 *>(slots + (size_t)s * slot_bytes + 32); };
    auto slot_flags = [&](int s) { return slots + (size_t)s * slot_bytes + 32 + (size_t)p.cap * 8; };

    // ---- CTA-wide, once: tap LUT (scipy.ndimage._gaussian_kernel1d) and barriers ---------------------------
    if (warp == 0) {
        double part = 0.0;
        for (int k = lane; k <= p.radius; k += 32)
            part += (k == 0 ? 1.0 : 2.0) * exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
            s_norm = part;
            for (int s = 0; s < RP_SLOTS; ++s) {
                mbar_init(&s_full[s], 32);  // every front lane arrives: each one releases its own writes to the slot
                mbar_init(&s_empty[s], RP_RENDER_WARPS * 32);  // every lane arrives: its own reads of the slot are released
                s_ticket[s] = 0;
            }
            for (int f = 0; f < NF; ++f) {
                mbar_init(&s_stage_all[f][0], 1);
                mbar_init(&s_stage_all[f][1], 1);
            }
            fence_mbar_init();
        }
    }
    __syncthreads();
    fill_lut(lut, p.n4, p.radius, p.sigma, 1.0 / s_norm, threadIdx.x, RN_THREADS);
    __syncthreads();

    if (warp < NF) {
        // =============================== front warp ==========================================================
        auto prefetch = [&](int t, int buf) {
            const uint32_t bx = (uint32_t)p.cap * 24u, bi = (uint32_t)p.cap * 8u;
            mbar_expect_tx(&s_stage[buf], bx + bi);
            bulk_g2s(stage[buf], p.xyz + (size_t)t * p.cap * 3, bx, &s_stage[buf]);
            bulk_g2s(stage[buf] + p.cap * 3, p.intensity + (size_t)t * p.cap, bi, &s_stage[buf]);
        };
        // templates are handed out by a global ticket (see render.cu); the front warp draws one ahead
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
        for (int j = 0;; ++j) {
            const int k = j * NF + front;  // global sequence number of this front's j-th template
            const int slot = k % RP_SLOTS, buf = j & 1;
            if (t >= p.n_tmpl) {  // out of work: hand the render warps a stop slot
                mbar_wait(&s_empty[slot], ((uint32_t)(k / RP_SLOTS) & 1u) ^ 1u);
                if (lane == 0) slot_header(slot)->t = -1;
                mbar_arrive(&s_full[slot]);
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
                mbar_wait(&s_stage[buf], (uint32_t)(j >> 1) & 1u);
                sxyz = stage[buf];
                sint = stage[buf] + p.cap * 3;
            }
            // the slot must have been drained by the render warps (first use passes immediately)
            mbar_wait(&s_empty[slot], ((uint32_t)(k / RP_SLOTS) & 1u) ^ 1u);
            PipeHeader *hd = slot_header(slot);
            uint2 *spots = slot_spots(slot);
            unsigned char *flags = slot_flags(slot);

            // ---- project, last-write-wins, compact (simulation2d.py:261-285, :422-430; detector_functions.py:297)
            for (int e = lane; e < p.table_size; e += 32) hash[e] = 0ull;
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                double px, py;
                project_spot(p, sxyz[3 * j], sxyz[3 * j + 1], px, py);
                int kk = -1;
                if (px >= 0.0 && px < (double)p.W && py >= 0.0 && py < (double)p.H) {
                    kk = (int)py * p.W + (int)px;
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
                const unsigned mask = __ballot_sync(0xffffffffu, live);
                if (live) {
                    const int d = n_live + __popc(mask & ((1u << lane) - 1u));
                    const int sx = kk % p.W, sy = kk / p.W;
                    const bool fold = sx < p.radius || sx >= p.W - p.radius || sy < p.radius || sy >= p.H - p.radius;
                    spots[d] = make_uint2((unsigned)sx | ((unsigned)(sy | (fold ? 0x4000 : 0)) << 16),
                                          __float_as_uint(inten[j]));
                }
                n_live += __popc(mask);
            }
            __syncwarp();

            // ---- normalisation: bound-pruned max pass (see render.cu) ---------------------------------------
            const int n_pass = (p.normalize && n_live > 0) ? 2 : 1;
            float scale = 1.f, vmax = INFINITY;
            if (n_pass == 2) {
                float amin = INFINITY, amax = -INFINITY;
                for (int j = lane; j < n_live; j += 32) {
                    amin = fminf(amin, spot_amp(spots[j]));
                    amax = fmaxf(amax, spot_amp(spots[j]));
                }
                amin = -warp_max(-amin);
                amax = warp_max(amax);
                const float w0 = tap(lut, p.n4, p.radius, 0);
                const float lower = amax * w0 * w0;
                const bool prune = amin >= 0.f && lower > 0.f;
                for (int reg0 = 0; reg0 < n_regions; reg0 += 32) {
                    const int reg = reg0 + lane;
                    if (reg < n_regions) {
                        unsigned char f = 1;
                        if (prune) {
                            const int rx0 = (reg % nrx) * RN_RW, ry0 = (reg / nrx) * RN_RH;
                            const int rx1 = min(rx0 + RN_RW, p.W) - 1, ry1 = min(ry0 + RN_RH, p.H) - 1;
                            float ub = 0.f;
                            for (int j = 0; j < n_live; ++j) {
                                const uint2 r = spots[j];
                                ub += spot_amp(r) * folded_bound(lut, p.n4, p.radius, rx0, rx1, spot_ix(r), p.W) *
                                      folded_bound(lut, p.n4, p.radius, ry0, ry1, spot_iy(r), p.H);
                            }
                            f = (ub * 1.001f >= lower) ? 1 : 0;
                        }
                        flags[reg] = f;
                    }
                }
                __syncwarp();
                FastSmem fs;
                fs.lut = lut;
                fs.spot = spots;
                float m = -INFINITY;
                for (int reg = 0; reg < n_regions; ++reg) {
                    if (!flags[reg]) continue;
                    const int rx0 = (reg % nrx) * RN_RW, ry0 = (reg / nrx) * RN_RH;
                    float acc[8][8];
                    bool mma;
                    if (!accumulate_region<false, DENSE>(p, fs, n_live, rx0, ry0, lane, acc, hits_s, mma)) {
                        m = fmaxf(m, 0.f);
                        continue;
                    }
                    m = fmaxf(m, region_max<VEC>(p, rx0, ry0, lane, acc, mma));
                }
                vmax = warp_max(m);
                scale = 1.f / vmax;  // np.divide(pattern, np.max(pattern)), simulation2d.py:440-441
            }
            if (lane == 0) {
                hd->n_live = n_live;
                hd->n_pass = n_pass;
                hd->t = t;
                hd->scale = scale;
                hd->vmax = vmax;
                s_ticket[slot] = 0;
            }
            __syncwarp();
            mbar_arrive(&s_full[slot]);  // release (all lanes): the slot is visible to the render warps
            t = t_next;
        }
    } else {
        // =============================== render warps ========================================================
        unsigned done = 0;  // bit f: front f has run out of work
        for (int k = 0; done != (1u << NF) - 1u; ++k) {
            if ((done >> (k % NF)) & 1u) continue;
            const int slot = k % RP_SLOTS;
            mbar_wait(&s_full[slot], (uint32_t)(k / RP_SLOTS) & 1u);
            const PipeHeader *hd = slot_header(slot);
            const int t = hd->t;
            if (t < 0) {
                done |= 1u << (k % NF);
                continue;
            }
            const int n_live = hd->n_live, n_pass = hd->n_pass;
            const float scale = hd->scale, vmax = hd->vmax;
            const unsigned char *flags = slot_flags(slot);
            FastSmem fs;
            fs.lut = lut;
            fs.spot = slot_spots(slot);
            float *img = p.images + (size_t)t * p.H * p.W;
            while (true) {
                int reg = 0;
                if (lane == 0) reg = atomicAdd(&s_ticket[slot], 1);
                reg = __shfl_sync(0xffffffffu, reg, 0);
                if (reg >= n_regions) break;
                const int rx0 = (reg % nrx) * RN_RW, ry0 = (reg / nrx) * RN_RH;
                float acc[8][8];
                bool mma;
                const bool any = accumulate_region<false, DENSE>(p, fs, n_live, rx0, ry0, lane, acc, hits_s, mma);
                if (!any && zero_tma) {
                    if (elect_one()) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap),
                                     "r"(smem_u32(s_zero)), "r"(rx0), "r"(ry0), "r"(t)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    continue;
                }
                float sc = any ? scale : 0.f;
                if (n_pass == 2 && any && flags[reg]) {  // pin the maximum pixel to exactly 1 (see render.cu)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = (acc[i][j] == vmax) ? 1.0f : acc[i][j] * sc;
                    sc = 1.0f;
                }
                store_region<VEC>(p, img, rx0, ry0, lane, acc, any, sc, mma);
            }
            mbar_arrive(&s_empty[slot]);
        }
        if (zero_tma && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    // the last CTA to finish re-arms the ticket counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&p.ticket[1], 1) == (int)gridDim.x - 1) {
        p.ticket[0] = 0;
        p.ticket[1] = 0;
    }
}

static int pipe_max_cap() {
    const int e = option(OPT_RENDER_PIPE_MAXCAP);
    return e >= 0 ? e : 512;  // beyond that (or when the slots do not fit in shared memory): render_kernel
}

// Returns 1 if the pipelined kernel was launched, 0 if the configuration is not eligible, < 0 on error.
int launch_render_pipelined(RenderParams p, cudaStream_t st) {
    const int n_regions = ((p.W + RN_RW - 1) / RN_RW) * ((p.H + RN_RH - 1) / RN_RH);
    if (p.cap > pipe_max_cap() || n_regions > 1024 || p.radius >= p.W || p.radius >= p.H) return 0;
    const int slot_bytes = (32 + p.cap * 8 + n_regions + 15) & ~15;
    const int front_bytes = (int)((p.stage ? (size_t)2 * p.cap * 32 : 0) + (size_t)p.table_size * 8 + (size_t)p.cap * 8);
    // Two front warps by default.  One keeps up only with the sparsest libraries (measured at 256 x 256, sigma 10,
    // normalised, capacity 32: 3.8 reflections per template 1258 us with one front vs 1273 with two, but 11.3
    // reflections 1611 vs 1277), so it takes the caller's word that the library is that sparse.
    int nf = (p.cap <= 32 && p.mean_spots_hint > 0.0 && p.mean_spots_hint < 6.0) ? 1 : 2;
    if (option(OPT_RENDER_FRONTS) > 0) nf = min(max(option(OPT_RENDER_FRONTS), 1), 3);
    const size_t smem = lut_smem_bytes(p.n4) + (size_t)p.hits_bytes + (size_t)nf * front_bytes + (size_t)2 * nf * slot_bytes;
    if (smem > 96 * 1024) return 0;
    const bool vec = (p.W & 3) == 0;
    // DENSE: the instantiation that carries the tensor-core path (larger code; only when hit lists were set up)
    void (*const kerns[2][3][2])(RenderParams, int, int, CUtensorMap, int, int) = {
        {{render_pipe_kernel<false, 1, false>, render_pipe_kernel<true, 1, false>},
         {render_pipe_kernel<false, 2, false>, render_pipe_kernel<true, 2, false>},
         {render_pipe_kernel<false, 3, false>, render_pipe_kernel<true, 3, false>}},
        {{render_pipe_kernel<false, 1, true>, render_pipe_kernel<true, 1, true>},
         {render_pipe_kernel<false, 2, true>, render_pipe_kernel<true, 2, true>},
         {render_pipe_kernel<false, 3, true>, render_pipe_kernel<true, 3, true>}}};
    const int dense = p.hits_bytes > 0 ? 1 : 0;
    void (*kern)(RenderParams, int, int, CUtensorMap, int, int) = kerns[dense][nf - 1][vec ? 1 : 0];
    // zero regions through the TMA engine: rows must be 16-byte multiples; the box is one warp region
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int zero_tma = 0;
    if (vec && option(OPT_RENDER_ZERO_TMA) > 0) {  // off by default: measured 4 % SLOWER on the headline (1312 vs 1262 us per 32 768 templates)
        const int rt = make_image_tensor_map(&tmap, p.images, p.n_tmpl, p.H, p.W, RN_RW, RN_RH, false);
        if (rt < 0) return rt;
        zero_tma = rt == 0 ? 1 : 0;
    }
    const int zero_tma_offset = (int)((smem + 127) & ~(size_t)127);
    const size_t smem_launch = zero_tma ? (size_t)zero_tma_offset + RN_RW * RN_RH * 4 : smem;
    if (smem_launch > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RN_THREADS, smem_launch);
    if (per_sm < 1) per_sm = 1;
    const int grid = p.n_tmpl < num_sms() * per_sm ? p.n_tmpl : num_sms() * per_sm;
    kern<<<grid, RN_THREADS, smem_launch, st>>>(p, slot_bytes, front_bytes, tmap, zero_tma, zero_tma_offset);
    const int rc = check_launch("ds_render (pipelined)");
    return rc == 0 ? 1 : rc;
}

}  // namespace ds
