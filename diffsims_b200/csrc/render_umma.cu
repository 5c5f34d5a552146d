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
//   * eight epilogue warps read the accumulators back with tcgen05.ld (lane = image row), take the template
//     maximum from them (no second evaluation), scale, write 32 x 32 tiles to 128-byte-swizzled shared memory and
//     hand them to the TMA engine (cp.async.bulk.tensor store through a 3-D tensor map of the image stack, which
//     also clips partial tiles), double-buffered so that tensor-memory reads, stores and copies overlap.
//
// Warp roles of the persistent CTA (one per SM, templates drawn from the global ticket):
//   warp 0        front: draws templates from the ticket and fetches their prepared records (render_prep.cu: live
//                 spots sorted by column, per-half lists) into one of four shared-memory slots with cp.async.bulk
//   warp 1        MMA issue (lane 0), owns the tensor-memory allocation (512 columns = 2 halves x 256)
//   warps 4..11   epilogue (tensor-memory lane quarter = warp % 4; two warps per quarter share its column tiles)
//   warps 2, 3, 12..15   operand producers: three teams of two warps, one shared-memory stage per team
// mbarriers: slot full (bulk copy -> everyone) / empty (everyone -> front), stage full/empty (producer <-> tcgen05.commit),
// half full/empty (tcgen05.commit <-> epilogue).  Bound: tensor pipe for dense templates, HBM write otherwise.
//
// Reference: diffsims/pattern/detector_functions.py:293-300 (assignment + scipy.ndimage.gaussian_filter),
// diffsims/simulations/simulation2d.py:261-285, :422-441.
#include <atomic>

#include "umma_device.cuh"

namespace ds {

constexpr int UM_NP = 3;                       // operand stages = producer teams
constexpr int UM_EPI = 8;                      // epilogue warps: two per tensor-memory lane quarter
// TEAM = producer warps per stage.  2: one warp per K group (spots 8 kh .. 8 kh + 7) writes A and B -- 16 warps, 128
// registers per thread.  4: per K group one warp writes A and another B -- 24 warps; the kernel starts with 80 registers
// per thread and setmaxnreg moves them: 128 for the two epilogue warpgroups, 56 for everyone else.
__host__ __device__ constexpr int um_warps(int team) { return team == 2 ? 16 : 24; }
// warp roles: 0 front, 1 MMA issue, 4-11 epilogue (lane quarter = warp % 4), producers 2-3 and 12.. (TEAM = 4: the last two
// warps of the CTA only fill their warpgroup)
constexpr int UM_EPI_WARP0 = 4;
constexpr int UM_A_BYTES = 128 * 16 * 2;       // one of A_hi / A_lo:  16 MN groups x 2 K groups x 128 B
constexpr int UM_B_BYTES = 256 * 16 * 2;       // one of B_hi / B_lo:  32 MN groups x 2 K groups x 128 B
constexpr int UM_STAGE_BYTES = 2 * UM_A_BYTES + 2 * UM_B_BYTES;  // 24 KB
constexpr int UM_TILE_BYTES = 32 * 32 * 4;     // one staged 32 x 32 float tile (128-byte rows, 128B swizzle)
constexpr int UM_BPAD = 8;                     // zero padding of the bf16 tap table on either side
constexpr int UM_MAX_CAP = 1024;
constexpr int UM_SLOTS = 4;                    // template records in flight per CTA

struct UmHeader {  // 32 bytes at the start of a record / slot (written by render_prepare_kernel)
    int t, n_live, n_half[2], pad[4];
};


// entries (16 bytes = 8 taps) per shifted copy of the bf16 tap table
__host__ __device__ inline int um_n8(int radius) { return (2 * radius + 2 * UM_BPAD + 1 + 7) / 8 + 1; }

// One lane's share of a chunk's A operand: four 16-byte units  a_s Wy_s[y]  for the rows 128 h + 8 (gq + 4 i) .. + 7.
// Straight-line code (the four units overlap): a unit the spot does not reach reads the all-zero head of the tap
// table.  scipy's mode="reflect" adds the mirror images of the spot at -c - 1 (reaches rows < R; LO halves) and at
// 2 H - 1 - c (rows >= H - R; HI halves), one more table read per unit where the half touches that border.
template <bool LO, bool HI>
__device__ __forceinline__ void produce_a(const LutRef &L, uint32_t a_hi, uint32_t a_lo, int gq, int row_base, bool ok, int cy,
                                          float am, int R, int H) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int g = gq + 4 * i, y_lo = row_base + 8 * g;
        const bool hit = ok && y_lo + 7 >= cy - R && y_lo <= cy + R;
        const uint32_t a = fetch_addr(L, hit ? y_lo + L.bias - cy : 0);
        float4 w0 = lds128(a), w1 = lds128(a + 16u);
        if (LO) {
            const int d = y_lo + cy + 1;
            const uint32_t a2 = fetch_addr(L, (hit && d <= R) ? d + L.bias : 0);
            w0 = add4(w0, lds128(a2));
            w1 = add4(w1, lds128(a2 + 16u));
        }
        if (HI) {
            const int d = y_lo + cy + 1 - 2 * H;
            const uint32_t a3 = fetch_addr(L, (hit && d + 7 >= -R && d <= R) ? d + L.bias : 0);
            w0 = add4(w0, lds128(a3));
            w1 = add4(w1, lds128(a3 + 16u));
        }
        store_split8(a_hi + 128u * g, a_lo + 128u * g, w0, w1, am);
    }
}
// One unit of the B operand in float32 (columns next to a border: direct taps plus the mirror images), then split.
__device__ __forceinline__ void produce_b_border(const LutRef &L, uint32_t hi_addr, uint32_t lo_addr, int x_lo, bool ok, int cx, int R,
                                                 int W) {
    const bool hit = ok && x_lo + 7 >= cx - R && x_lo <= cx + R;
    const uint32_t a1 = fetch_addr(L, hit ? x_lo + L.bias - cx : 0);
    float4 w0 = lds128(a1), w1 = lds128(a1 + 16u);
    const int d2 = x_lo + cx + 1, d3 = x_lo + cx + 1 - 2 * W;
    const uint32_t a2 = fetch_addr(L, (hit && d2 <= R) ? d2 + L.bias : 0);
    const uint32_t a3 = fetch_addr(L, (hit && d3 + 7 >= -R && d3 <= R) ? d3 + L.bias : 0);
    w0 = add4(add4(w0, lds128(a2)), lds128(a3));
    w1 = add4(add4(w1, lds128(a2 + 16u)), lds128(a3 + 16u));
    store_split8(hi_addr, lo_addr, w0, w1, 1.0f);
}

#ifdef DS_PROF
// per-role cycle counters of CTA 0 (profiling builds only): [role][0] = cycles in the role's loop, [1] = of which waiting
__device__ unsigned long long g_um_prof[4][2];
__device__ unsigned long long g_um_prof2[8];  // first epilogue warp of CTA 0: cycles per section (register accumulators)
#define PROF_SEC_DECL long long prof_sec[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_s0 = 0
#define PROF_SEC_BEGIN prof_s0 = clock64()
#define PROF_SEC(i)                        \
    {                                      \
        const long long now_ = clock64();  \
        prof_sec[i] += now_ - prof_s0;     \
        prof_s0 = now_;                    \
    }
#define PROF_SEC_STORE(first, count)                                                 \
    if (blockIdx.x == 0 && lane == 0)                                                \
        for (int i_ = first; i_ < first + count; ++i_) g_um_prof2[i_] = (unsigned long long)prof_sec[i_];
#define PROF_DECL long long prof_t0 = clock64(), prof_wait = 0, prof_w0 = 0
#define PROF_WAIT_BEGIN prof_w0 = clock64()
#define PROF_WAIT_END prof_wait += clock64() - prof_w0
#define PROF_DONE(role)                                                          \
    if (blockIdx.x == 0 && lane == 0) {                                          \
        g_um_prof[role][0] = (unsigned long long)(clock64() - prof_t0);          \
        g_um_prof[role][1] = (unsigned long long)prof_wait;                      \
    }
#else
#define PROF_SEC_DECL
#define PROF_SEC_BEGIN
#define PROF_SEC(i)
#define PROF_SEC_STORE(first, count)
#define PROF_DECL
#define PROF_WAIT_BEGIN
#define PROF_WAIT_END
#define PROF_DONE(role)
#endif

template <int TEAM>
__global__ void __launch_bounds__(um_warps(TEAM) * 32, 1) render_umma_kernel(const RenderParams p, const __grid_constant__ CUtensorMap tmap,
                                                                              const unsigned char *records, const int slot_bytes,
                                                                              const int window, const int epi_bufs) {
    constexpr int UM_PROD = TEAM * UM_NP;            // producer warps
    constexpr int UM_THREADS = um_warps(TEAM) * 32;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_slot_full[UM_SLOTS], s_slot_empty[UM_SLOTS], s_stage_full[UM_NP], s_stage_empty[UM_NP],
        s_half_full[2], s_half_empty[2];
    __shared__ float s_emax[2][UM_EPI];
    __shared__ uint32_t s_tmem;
    __shared__ double s_norm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.radius, H = p.H, W = p.W;
    const int n_halves = (H + 127) >> 7;
    const int Wp = (W + 15) & ~15;  // columns of a full-width product
    const int n8 = um_n8(R);

    // ---- shared memory: stages | epilogue tiles (1024-byte aligned) | tap LUT (float, 4 shifted copies) | bf16 tap
    // table (8 shifted copies, hi then lo) | 4 slots ---------------------------------------------------------------
    unsigned char *stages = smem_raw;
    unsigned char *epi = stages + (size_t)UM_NP * UM_STAGE_BYTES;
    float4 *lut = reinterpret_cast<float4 *>(epi + (size_t)UM_EPI * epi_bufs * UM_TILE_BYTES);
    unsigned char *btab = reinterpret_cast<unsigned char *>(lut) + lut_smem_bytes(p.n4);
    unsigned char *slots = btab + (size_t)2 * 8 * n8 * 16;
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
            for (int s = 0; s < UM_SLOTS; ++s) {
                mbar_init(&s_slot_full[s], 1);
                mbar_init(&s_slot_empty[s], 1 + UM_EPI + UM_PROD);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&s_half_full[s], 1);
                mbar_init(&s_half_empty[s], UM_EPI);
            }
            for (int s = 0; s < UM_NP; ++s) {
                mbar_init(&s_stage_full[s], TEAM);
                mbar_init(&s_stage_empty[s], 1);
            }
            fence_mbar_init();
        }
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp == UM_EPI_WARP0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
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
    // Role dispatch.  With TEAM = 4 the two epilogue warpgroups (warps 4..11) raise their register allowance and every
    // other warpgroup lowers it; each setmaxnreg sits at the head of the branch it governs (one instruction per
    // warpgroup, and ptxas allocates each branch against its own limit).
    const bool epi_role = warp >= UM_EPI_WARP0 && warp < UM_EPI_WARP0 + UM_EPI;
    if (!epi_role) {
    if (TEAM == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // =============================== front warp ==============================================================
        PROF_DECL;
        // The ticket atomic is a global round trip of a microsecond behind a saturated store stream, so it runs one
        // template ahead of the record fetch (which is itself up to UM_SLOTS templates ahead of the consumers).
        int t_ahead = 0;
        if (lane == 0) t_ahead = atomicAdd(&p.ticket[0], 1);
        for (int k = 0;; ++k) {
            const int slot = k % UM_SLOTS;
            const int t = __shfl_sync(0xffffffffu, t_ahead, 0);
            if (lane == 0 && t < p.n_tmpl) t_ahead = atomicAdd(&p.ticket[0], 1);
            PROF_WAIT_BEGIN;
            mbar_wait(&s_slot_empty[slot], ((uint32_t)(k / UM_SLOTS) & 1u) ^ 1u);
            PROF_WAIT_END;
            if (t >= p.n_tmpl) {  // out of work: stop slot
                if (lane == 0) {
                    slot_header(slot)->t = -1;
                    mbar_arrive(&s_slot_full[slot]);
                }
                break;
            }
            if (lane == 0) {  // the copy's completion is the slot's `full` signal
                mbar_expect_tx(&s_slot_full[slot], (uint32_t)slot_bytes);
                bulk_g2s(slots + (size_t)slot * slot_bytes, records + (size_t)t * slot_bytes, (uint32_t)slot_bytes, &s_slot_full[slot]);
            }
            __syncwarp();
        }
        PROF_DONE(0);
    } else if (warp == 1) {
        // =============================== MMA issue ===============================================================
        PROF_DECL;
        PROF_SEC_DECL;
        int stage = 0;           // chunks go round the UM_NP stages; `sphase` is the parity of a stage's current fill
        uint32_t sphase = 0;
        const uint32_t st0 = smem_u32(stages);
        // shared-memory descriptors of stage 0; a stage further on adds its byte offset >> 4 to the address field
        const uint64_t d_a_hi = umma_desc(st0, 16 * 128, 128), d_a_lo = umma_desc(st0 + UM_A_BYTES, 16 * 128, 128);
        const uint64_t d_b_hi = umma_desc(st0 + 2 * UM_A_BYTES, 32 * 128, 128);
        const uint64_t d_b_lo = umma_desc(st0 + 2 * UM_A_BYTES + UM_B_BYTES, 32 * 128, 128);
        for (int k = 0;; ++k) {
            const int slot = k % UM_SLOTS;
            PROF_WAIT_BEGIN;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k / UM_SLOTS) & 1u);
            PROF_WAIT_END;
            const UmHeader *hd = slot_header(slot);
            if (hd->t < 0) break;
            const uint32_t *wins = reinterpret_cast<const uint32_t *>(slots + (size_t)slot * slot_bytes + umma_windows_offset(p.cap));
            for (int h = 0; h < n_halves; ++h) {
                const int n_chunks = max(1, (hd->n_half[h] + 15) >> 4);
                PROF_WAIT_BEGIN;
                mbar_wait(&s_half_empty[h], ((uint32_t)k & 1u) ^ 1u);  // the epilogue has drained template k - 1
                PROF_WAIT_END;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t *win_h = wins + h * ((p.cap + 15) >> 4);
                    for (int j = 0; j < n_chunks; ++j) {
                        PROF_WAIT_BEGIN;
                        mbar_wait(&s_stage_full[stage], sphase);
                        PROF_WAIT_END;
                        PROF_SEC_BEGIN;
                        tc_fence_after();
                        const uint32_t win = win_h[j];
                        const uint64_t so = (uint64_t)((uint32_t)stage * (UM_STAGE_BYTES >> 4));
                        const uint32_t idesc = umma_idesc((int)(win >> 16));
                        const uint32_t d = tm + (uint32_t)(h * 256) + (win & 0xffffu);
                        PROF_SEC(0);
                        umma_keep_a(d, d_a_hi + so, d_b_hi + so, idesc, j > 0);
                        umma_reuse_a(d, d_a_hi + so, d_b_lo + so, idesc, 1);
                        umma(d, d_a_lo + so, d_b_hi + so, idesc, 1);
                        umma_commit(&s_stage_empty[stage]);                   // the stage may be refilled
                        if (j == n_chunks - 1) umma_commit(&s_half_full[h]);  // the half is complete
                        PROF_SEC(1);
                        if (++stage == UM_NP) {
                            stage = 0;
                            sphase ^= 1u;
                        }
                    }
                }
                // (the other lanes only follow the template / half structure)
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
        PROF_DONE(1);
        PROF_SEC_STORE(0, 2);
    } else if ((warp < UM_EPI_WARP0 ? warp - 2 : warp - (UM_EPI_WARP0 + UM_EPI) + 2) < UM_PROD) {
        // =============================== operand producers =======================================================
        PROF_DECL;
        PROF_SEC_DECL;
        const int pi = warp < UM_EPI_WARP0 ? warp - 2 : warp - (UM_EPI_WARP0 + UM_EPI) + 2;  // 0 .. UM_PROD - 1
        // team (= its stage), the K group (spots 8 kh .. 8 kh + 7) this warp writes, and which operand(s)
        const int pw = pi / TEAM, kh = pi & 1;
        const bool do_a = TEAM == 2 || ((pi % TEAM) >> 1) == 0, do_b = TEAM == 2 || ((pi % TEAM) >> 1) == 1;
        const uint32_t st = smem_u32(stages) + (uint32_t)pw * UM_STAGE_BYTES;
        const int k8 = lane & 7, gq = lane >> 3;
        LutRef L;
        L.base = smem_u32(lut);
        L.n4 = p.n4;
        L.last = p.n4 - 1;
        L.bias = R + LUT_PAD;
        const uint32_t bt_hi = smem_u32(btab), bt_lo = bt_hi + (uint32_t)(8 * n8 * 16);
        const uint32_t a_hi = st + (uint32_t)(kh * (16 * 128) + k8 * 16), a_lo = a_hi + UM_A_BYTES;
        const uint32_t b_hi = st + 2 * UM_A_BYTES + (uint32_t)(kh * (32 * 128) + k8 * 16), b_lo = b_hi + UM_B_BYTES;
        int c = 0;
        for (int k = 0;; ++k) {
            const int slot = k % UM_SLOTS;
            PROF_WAIT_BEGIN;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k / UM_SLOTS) & 1u);
            PROF_WAIT_END;
            const UmHeader *hd = slot_header(slot);
            if (hd->t < 0) break;
            const uint32_t spot_s = smem_u32(slot_spots(slot));
            for (int h = 0; h < n_halves; ++h) {
                const int n_h = hd->n_half[h];
                const int n_chunks = max(1, (n_h + 15) >> 4);
                const uint32_t list_s = smem_u32(slot_list(slot, h));
                const uint32_t win_s = smem_u32(slots + (size_t)slot * slot_bytes + umma_windows_offset(p.cap));
                const int win_stride = (p.cap + 15) >> 4;
                for (int j = 0; j < n_chunks; ++j, ++c) {
                    if (c % UM_NP != pw) continue;
                    PROF_SEC_BEGIN;
                    // this lane's spot of the chunk and the chunk's column window (render_prep.cu)
                    const int li = 16 * j + 8 * kh + k8;
                    const bool ok = li < n_h;
                    int cx = 0, cy = 0;
                    float am = 0.f;
                    if (ok) {
                        const uint2 r = lds64v(spot_s + 8u * lds16(list_s + 2u * (uint32_t)li));
                        cx = (int)(r.x & 0xffffu);
                        cy = (int)(r.x >> 16);
                        am = __uint_as_float(r.y);
                    }
                    const uint32_t win = lds32v(win_s + 4u * (uint32_t)(h * win_stride + j));
                    const int col0 = (int)(win & 0xffffu), ncols = (int)(win >> 16);
                    PROF_SEC(2);
                    PROF_WAIT_BEGIN;
                    mbar_wait(&s_stage_empty[pw], ((uint32_t)(c / UM_NP) & 1u) ^ 1u);
                    PROF_WAIT_END;
                    PROF_SEC_BEGIN;
                    // ---- A
                    const bool lo_half = 128 * h < R, hi_half = 128 * h + 127 >= H - R;
                    if (!do_a) {
                    } else if (lo_half && hi_half)
                        produce_a<true, true>(L, a_hi, a_lo, gq, 128 * h, ok, cy, am, R, H);
                    else if (lo_half)
                        produce_a<true, false>(L, a_hi, a_lo, gq, 128 * h, ok, cy, am, R, H);
                    else if (hi_half)
                        produce_a<false, true>(L, a_hi, a_lo, gq, 128 * h, ok, cy, am, R, H);
                    else
                        produce_a<false, false>(L, a_hi, a_lo, gq, 128 * h, ok, cy, am, R, H);
                    PROF_SEC(3);
                    // ---- B: Wx_s[x], columns col0 + 8 g .. + 7.  Blocks of four units next to a border take the float32
                    // path with the mirror images; everything in between is a shifted copy of the pre-split table
                    const int n_blk = ncols >> 5, n_units = do_b ? ncols >> 3 : 0;  // (ncols is a multiple of 16: the last block may be half)
                    int g0 = do_b ? 0 : ncols;
#pragma unroll 1
                    for (; 8 * g0 < ncols && col0 + 8 * g0 < R; g0 += 4)
                        if (g0 + gq < n_units) produce_b_border(L, b_hi + 128u * (g0 + gq), b_lo + 128u * (g0 + gq), col0 + 8 * (g0 + gq), ok, cx, R, W);
#pragma unroll 2
                    for (; 8 * g0 < ncols && col0 + 8 * g0 + 31 < W - R; g0 += 4) {
                        const int g = g0 + gq, x_lo = col0 + 8 * g;
                        const bool hit = ok && x_lo + 7 >= cx - R && x_lo <= cx + R;
                        const int a = hit ? x_lo - cx + R + UM_BPAD : 0;  // (entry 0 of copy 0 is all zero)
                        const uint32_t off = (uint32_t)(((a & 7) * n8 + (a >> 3)) << 4);
                        const uint4 vh = lds128u(bt_hi + off), vl = lds128u(bt_lo + off);
                        if (g < n_units) {
                            sts128(b_hi + 128u * g, vh.x, vh.y, vh.z, vh.w);
                            sts128(b_lo + 128u * g, vl.x, vl.y, vl.z, vl.w);
                        }
                    }
#pragma unroll 1
                    for (; 8 * g0 < ncols; g0 += 4)
                        if (g0 + gq < n_units) produce_b_border(L, b_hi + 128u * (g0 + gq), b_lo + 128u * (g0 + gq), col0 + 8 * (g0 + gq), ok, cx, R, W);
                    (void)n_blk;
                    PROF_SEC(4);
                    proxy_fence();  // the stores above become visible to the tensor core's (async proxy) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_stage_full[pw]);
                    PROF_SEC(5);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_slot_empty[slot]);
        }
        if (pi == 0) {
            PROF_DONE(3);
            PROF_SEC_STORE(2, 4);
        }
    }
    } else {
        if (TEAM == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        // =============================== epilogue ================================================================
        PROF_DECL;
        const int ew = warp - UM_EPI_WARP0;  // 0..7
        const int q = warp & 3;              // tensor-memory lane quarter this warp may read
        const int ch = ew >> 2;              // of the 32-column tiles of a half this warp takes ct = ch, ch + 2, ...
        const uint32_t tile_s = smem_u32(epi + (size_t)ew * epi_bufs * UM_TILE_BYTES);
        const int n_ct = (W + 31) >> 5;
        const uint32_t tm_q = tm + ((uint32_t)(32 * q) << 16);
        int nbuf = 0;  // staged tiles so far (buffer = nbuf & 1)
        for (int k = 0;; ++k) {
            const int slot = k % UM_SLOTS;
            PROF_WAIT_BEGIN;
            mbar_wait(&s_slot_full[slot], (uint32_t)(k / UM_SLOTS) & 1u);
            PROF_WAIT_END;
            const UmHeader *hd = slot_header(slot);
            const int t = hd->t;
            if (t < 0) break;
            const bool norm = p.normalize && hd->n_live > 0;  // no spot in frame: zeros, returned un-normalised
            float scale = 1.f, vmax = INFINITY, clamp = __uint_as_float(0x7fc00000u);
            if (norm) {
                // ---- pass 1: the template maximum, straight from the accumulators.  (Loops are kept rolled: the four
                // roles of this kernel share the instruction cache.)
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
                for (int h = 0; h < n_halves; ++h) {
                    PROF_WAIT_BEGIN;
                    mbar_wait(&s_half_full[h], (uint32_t)k & 1u);
                    PROF_WAIT_END;
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
                PROF_WAIT_BEGIN;
                asm volatile("bar.sync 1, %0;" ::"n"(UM_EPI * 32) : "memory");
                PROF_WAIT_END;
#pragma unroll
                for (int e = 0; e < UM_EPI; ++e) m = fmaxf(m, s_emax[k & 1][e]);
                vmax = m;
                // np.divide(pattern, np.max(pattern)), simulation2d.py:440-441: the maximum pixel is exactly 1 there.
                // x * (1 / max) can be 1 ulp off, so the reciprocal is rounded UP (max * scale >= 1 in exact arithmetic,
                // hence after rounding) and the product clamped to 1.  A NaN clamp leaves everything as computed when
                // the maximum is not a positive finite number (fminf returns the other operand).
                scale = 1.f / vmax;
                if (vmax > 0.f && vmax < INFINITY) {
                    if (scale * vmax < 1.f) scale = __uint_as_float(__float_as_uint(scale) + 1u);
                    clamp = 1.f;
                }
            }
            // ---- pass 2: scale, stage, store.  Two register sets alternate (the load of the next tile is in flight
            // while this one is scaled and staged); `stage_tile` is the per-tile tail.
            auto stage_tile = [&](const uint32_t (&r)[32], int ct, int row0) {
                // the copy that last read this buffer has finished reading it
                if (elect_one()) {
                    if (epi_bufs == 2)
                        tma_store_wait_read<1>();
                    else
                        tma_store_wait_read<0>();
                }
                __syncwarp();
                const uint32_t buf = tile_s + (uint32_t)(epi_bufs == 2 ? (nbuf & 1) : 0) * UM_TILE_BYTES;
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
                    PROF_WAIT_BEGIN;
                    mbar_wait(&s_half_full[h], (uint32_t)k & 1u);
                    PROF_WAIT_END;
                    tc_fence_after();
                }
                const int row0 = 128 * h + 32 * q;  // first image row of this warp's 32 lanes
                if (row0 < H && ch < n_ct) {
                    uint32_t ra[32], rb[32];
                    tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * ch), ra);
#pragma unroll 1
                    for (int ct = ch; ct < n_ct; ct += 4) {
                        if (ct + 2 < n_ct) tmem_ld32_issue(tm_q + (uint32_t)(h * 256 + 32 * (ct + 2)), rb);
                        tmem_wait(ra);  // (all loads issued so far: the one just issued is short)
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
        if (ew == 0) { PROF_DONE(2); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart statically and has no
// link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static std::atomic<void *> cached{nullptr};
    void *fn = cached.load(std::memory_order_acquire);
    if (fn == nullptr) {
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            fn = nullptr;
        cudaGetLastError();
        cached.store(fn, std::memory_order_release);
    }
    return reinterpret_cast<EncodeTiledFn>(fn);
}

// The image stack as a 3-D tensor (W, H, n_tmpl) of float32 for cp.async.bulk.tensor stores; a box is box_w x box_h pixels of
// one template (partial boxes at the image edges are clipped by the TMA engine).  0 = ok, 1 = the driver entry point is not
// available (callers fall back to st.global), < 0 = error.
int make_image_tensor_map(CUtensorMap *tmap, float *images, int n_tmpl, int H, int W, int box_w, int box_h, bool swizzle128) {
    EncodeTiledFn encode = encode_tiled();
    if (encode == nullptr) return 1;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_tmpl};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1}, estr[3] = {1, 1, 1};
    const CUresult cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, images, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("ds_render: cuTensorMapEncodeTiled failed (%d)", (int)cr);
        return -2;
    }
    return 0;
}

int launch_render_prepare(const RenderParams &p, unsigned char *records, int window, cudaStream_t st);

// Returns 1 if the tensor-core kernel was launched, 0 if the configuration is not eligible, < 0 on error.
// `records`: n_tmpl * umma_record_bytes(cap) bytes of device scratch (the prepared templates).
bool umma_eligible(int H, int W, int cap, int radius) {
    return !(H > 256 || W > 256 || (W & 3) != 0 || cap > UM_MAX_CAP || radius >= W || radius >= H || radius > 120);
}

int launch_render_umma(RenderParams p, unsigned char *records, cudaStream_t st) {
    if (!umma_eligible(p.H, p.W, p.cap, p.radius)) return 0;
    const int n8 = um_n8(p.radius);
    const int slot_bytes = umma_record_bytes(p.cap);
    // two staged tiles per epilogue warp when they fit (large capacities are tensor-pipe bound anyway)
    int epi_bufs = 2;
    auto smem_for = [&](int bufs) {
        return (size_t)UM_NP * UM_STAGE_BYTES + (size_t)UM_EPI * bufs * UM_TILE_BYTES + lut_smem_bytes(p.n4) + (size_t)2 * 8 * n8 * 16 +
               (size_t)UM_SLOTS * slot_bytes;
    };
    if (smem_for(2) > 226 * 1024) epi_bufs = 1;
    const size_t smem = smem_for(epi_bufs);
    if (smem > 226 * 1024) return 0;  // (+ ~1 KB of static shared memory: barriers, alignment)
    alignas(64) CUtensorMap tmap;
    const int rt = make_image_tensor_map(&tmap, p.images, p.n_tmpl, p.H, p.W, 32, 32, true);
    if (rt != 0) return rt < 0 ? rt : 0;
    const int window = option(OPT_RENDER_UMMA_WINDOW) == 0 ? 0 : 1;
    const int rc0 = launch_render_prepare(p, records, window, st);
    if (rc0 != 0) return rc0;
    const int sms = num_sms();
    const int grid = p.n_tmpl < sms ? p.n_tmpl : sms;
    // producer warps per operand stage: 2; render_umma_team = 4 puts A and B of a K group on separate warps (24 warps,
    // setmaxnreg) -- measured 3 - 6 % slower on every configuration (the shared-memory pipe, not producer latency, is the
    // limit), kept for the A/B measurement
    if (option(OPT_RENDER_UMMA_TEAM) != 4) {
        cudaFuncSetAttribute(render_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        render_umma_kernel<2><<<grid, um_warps(2) * 32, smem, st>>>(p, tmap, records, slot_bytes, window, epi_bufs);
    } else {
        cudaFuncSetAttribute(render_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        render_umma_kernel<4><<<grid, um_warps(4) * 32, smem, st>>>(p, tmap, records, slot_bytes, window, epi_bufs);
    }
    const int rc = check_launch("ds_render (tcgen05)");
    return rc == 0 ? 1 : rc;
}

}  // namespace ds

#ifdef DS_PROF
extern "C" int ds_debug_umma_prof(unsigned long long *out_host /*[16]*/) {
    if (cudaMemcpyFromSymbol(out_host, ds::g_um_prof, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(out_host + 8, ds::g_um_prof2, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : -1;
}
#endif
