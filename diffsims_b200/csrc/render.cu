// K3 -- rasterise per-rotation spot lists into templates (float32 images).
//
// Replaces, per template:
//   Simulation2D._get_transformed_coordinates   diffsims/simulations/simulation2d.py:261-285
//   get_diffraction_pattern                      simulation2d.py:357-442  (in-frame, astype(int), /max)
//   get_pattern_from_pixel_coordinates_and_intensities  diffsims/pattern/detector_functions.py:251-311
//     fast:  out[y, x] = I (assignment, last write wins) ; scipy.ndimage.gaussian_filter(out, sigma)
//     slow:  _subpixel_gaussian (:314-359), additive clipped boxes
//
// The reference blurs a full H x W image with two 1-D passes.  Here the blur is applied analytically: the
// image of a delta at pixel (ix, iy) under a separable, reflect-folded kernel is the outer product
//   Wy(y; iy) * Wx(x; ix),  Wx(x; ix) = w[x - ix] + w[x + ix + 1] + w[x + ix + 1 - 2 W]
// (direct tap + the two mirror images of scipy's mode="reflect"), so every pixel is a gather over the few
// spots whose (2r+1)^2 box reaches it.  One CTA (8 warps) owns one template.  A warp owns 64 x 32 pixel
// regions; a lane owns an 8 x 8 register tile (two float4 column groups 32 px apart, so every warp-wide
// float4 store writes 4 rows x 128 contiguous bytes).  Spots are culled per warp region with a ballot;
// weights come from a zero-padded, 4-way shifted LUT in shared memory (aligned LDS.128 for any offset).
// Output is written exactly once with coalesced 16-byte stores: the kernel's algorithmic traffic is
// H*W*4 bytes per template and it is HBM-write bound; normalisation by the template maximum is done by
// recomputing the register tiles in a second pass instead of re-reading the image.
#include "render_device.cuh"

namespace ds {

int launch_render_pipelined(RenderParams p, cudaStream_t st);
int launch_render_umma(RenderParams p, unsigned char *records, cudaStream_t st);
int launch_render_rows(RenderParams p, unsigned char *records, cudaStream_t st);
bool umma_eligible(int H, int W, int cap, int radius);
// the dispatch rule of ds_render for the tcgen05 path (shared with ds_render_launch_count)
static bool wants_umma(int cap, double mean_spots_hint) {
    const int um = option(OPT_RENDER_UMMA);
    return um >= 0 ? um != 0 : (mean_spots_hint > 0.0 ? mean_spots_hint >= 16.0 : cap >= 96);
}

// A "group" of G warps (G = 1, 2, 4 or 8) owns one template at a time; a CTA holds 8 / G groups and is
// persistent (templates are drawn from a global ticket).  G = 1 keeps 8 independent templates in flight per
// CTA with no block-wide barriers at all (sparse patterns are latency-, not throughput-bound); larger G
// spreads the regions of one dense template over more warps.  Groups synchronise on named barriers.
template <int G>
__device__ __forceinline__ void group_sync(int group) {
    if (G == 1) {
        __syncwarp();
    } else if (G == RN_WARPS) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(G * 32) : "memory");
    }
}

template <bool FAST, int G, bool WIDE, bool VEC, bool DENSE>
__global__ void __launch_bounds__(RN_THREADS, 2) render_kernel(const RenderParams p, const int group_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NGROUPS = RN_WARPS / G;
    constexpr int GT = G * 32;  // threads per group
    __shared__ int s_n_live[NGROUPS], s_n_inframe[NGROUPS];
    __shared__ float s_wmax[RN_WARPS];
    __shared__ double s_norm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / G, gwarp = warp % G, gtid = gwarp * 32 + lane;

    // ---- CTA-wide, once: the 4-way shifted LUT of the normalised 1-D taps --------------------------------
    // scipy.ndimage._gaussian_kernel1d: exp(-0.5 k^2 / sigma^2) / sum over |k| <= radius
    float4 *lut = reinterpret_cast<float4 *>(smem_raw);
    unsigned char *gbase =
        smem_raw + (FAST ? lut_smem_bytes(p.n4) + (size_t)p.hits_bytes : 0) + (size_t)group * group_bytes;
    const uint32_t hits_s = (FAST && p.hits_bytes)
                                ? smem_u32(smem_raw + lut_smem_bytes(p.n4)) +
                                      (uint32_t)(warp * (p.hits_bytes / RN_WARPS))
                                : 0u;
    if (FAST) {
        if (warp == 0) {
            double part = 0.0;
            for (int k = lane; k <= p.radius; k += 32)
                part += (k == 0 ? 1.0 : 2.0) * exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_norm = part;
        }
        __syncthreads();
        fill_lut(lut, p.n4, p.radius, p.sigma, 1.0 / s_norm, threadIdx.x, RN_THREADS);
        __syncthreads();
    }
    const int nrx = (p.W + RN_RW - 1) / RN_RW, nry = (p.H + RN_RH - 1) / RN_RH;
    const int n_regions = nrx * nry;
    FastSmem fs;
    SlowSmem ss;
    unsigned char *flags;  // [n_regions] 1 = the region can hold the template maximum
    float *ubound;         // [G][n_regions] partial upper bounds
    // double-buffered staging of the next template's spot rows (xyz [cap][3] + intensity [cap], float64),
    // filled by cp.async.bulk one template ahead: under a saturated write stream a dependent global load
    // costs microseconds, which would otherwise serialise count -> spots -> pixels for every template
    double *stage[2] = {nullptr, nullptr};
    const int stage_elems = p.cap * 4;
    if (p.stage) {
        stage[0] = reinterpret_cast<double *>(gbase);
        stage[1] = stage[0] + stage_elems;
        gbase += (size_t)2 * stage_elems * sizeof(double);
    }
    if (FAST) {
        fs = carve_fast(gbase, p);
        fs.lut = lut;
        flags = reinterpret_cast<unsigned char *>(fs.spot + p.cap);
        ubound = reinterpret_cast<float *>(flags + ((n_regions + 3) & ~3));
    } else {
        ss = carve_slow(gbase, p);
        flags = reinterpret_cast<unsigned char *>(ss.yhi + p.cap);
        ubound = reinterpret_cast<float *>(flags + ((n_regions + 3) & ~3));
    }
    __shared__ __align__(8) uint64_t s_bar[NGROUPS][2];
    if (p.stage) {
        if (gtid == 0) {
            mbar_init(&s_bar[group][0], 1);
            mbar_init(&s_bar[group][1], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    auto prefetch = [&](int t, int buf) {  // called by one thread of the group
        const uint32_t bx = (uint32_t)p.cap * 24u, bi = (uint32_t)p.cap * 8u;
        mbar_expect_tx(&s_bar[group][buf], bx + bi);
        bulk_g2s(stage[buf], p.xyz + (size_t)t * p.cap * 3, bx, &s_bar[group][buf]);
        bulk_g2s(stage[buf] + p.cap * 3, p.intensity + (size_t)t * p.cap, bi, &s_bar[group][buf]);
    };

    // Templates are handed out by a global ticket, not by a static grid stride: per-SM write bandwidth is
    // not uniform on B200 and a static split leaves the fast SMs idle at the end (measured with
    // tools/microbench/write_bw.cu: 6.3 TB/s static vs 7.4 TB/s ticketed for the same store pattern).
    __shared__ int s_draw[NGROUPS];
    auto draw = [&]() {  // one ticket per group, broadcast through shared memory
        if (gtid == 0) s_draw[group] = atomicAdd(&p.ticket[0], 1);
        group_sync<G>(group);
        const int t = s_draw[group];
        group_sync<G>(group);
        return t;
    };
    int t = draw();
    int n_next = 0, iter = 0;
    if (t < p.n_tmpl) {
        n_next = p.count[t];
        if (p.stage && gtid == 0) prefetch(t, 0);
    }
    for (; t < p.n_tmpl; ++iter) {
        // ---- project spots to detector pixels (simulation2d.py:261-285, :422-430), float64 -----------
        const int n = min(n_next, p.cap);
        const int buf = iter & 1;
        const int t_next = draw();
        if (t_next < p.n_tmpl) {  // one template ahead: count in a register, spot rows by bulk copy
            n_next = p.count[t_next];
            if (p.stage && gtid == 0) prefetch(t_next, buf ^ 1);
        }
        const double *sxyz = p.xyz + (size_t)t * p.cap * 3;
        const double *sint = p.intensity + (size_t)t * p.cap;
        if (p.stage) {
            mbar_wait(&s_bar[group][buf], (uint32_t)(iter >> 1) & 1u);
            sxyz = stage[buf];
            sint = stage[buf] + p.cap * 3;
        }
        if (FAST) {
            for (int e = gtid; e < p.table_size; e += GT) fs.hash[e] = 0ull;
            group_sync<G>(group);
            for (int j = gtid; j < n; j += GT) {
                double px, py;
                project_spot(p, sxyz[3 * j], sxyz[3 * j + 1], px, py);
                int key = -1;
                if (px >= 0.0 && px < (double)p.W && py >= 0.0 && py < (double)p.H) {
                    key = (int)py * p.W + (int)px;  // astype(int): truncation
                    // last write wins (detector_functions.py:297): keep the largest spot index per pixel
                    const unsigned long long packed = ((unsigned long long)(key + 1) << 32) | (unsigned)j;
                    unsigned h = ((unsigned)key * 2654435761u) & (p.table_size - 1);
                    while (true) {
                        unsigned long long cur = fs.hash[h];
                        if (cur == 0ull) {
                            const unsigned long long old = atomicCAS(&fs.hash[h], 0ull, packed);
                            if (old == 0ull) break;
                            cur = old;
                        }
                        if ((cur >> 32) == (unsigned long long)(key + 1)) {
                            atomicMax(&fs.hash[h], packed);
                            break;
                        }
                        h = (h + 1) & (p.table_size - 1);
                    }
                }
                fs.key[j] = key;
                fs.inten[j] = (float)sint[j];
            }
            group_sync<G>(group);
            if (gwarp == 0) {  // deterministic compaction of the surviving spots
                int n_live = 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    bool live = false;
                    int key = -1;
                    if (j < n && (key = fs.key[j]) >= 0) {
                        unsigned h = ((unsigned)key * 2654435761u) & (p.table_size - 1);
                        while ((fs.hash[h] >> 32) != (unsigned long long)(key + 1)) h = (h + 1) & (p.table_size - 1);
                        live = (unsigned)(fs.hash[h] & 0xffffffffu) == (unsigned)j;
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, live);
                    if (live) {
                        const int d = n_live + __popc(mask & ((1u << lane) - 1u));
                        const int sx = key % p.W, sy = key / p.W;
                        const bool fold = sx < p.radius || sx >= p.W - p.radius || sy < p.radius || sy >= p.H - p.radius;
                        fs.spot[d] = make_uint2((unsigned)sx | ((unsigned)(sy | (fold ? 0x4000 : 0)) << 16),
                                                __float_as_uint(fs.inten[j]));
                    }
                    n_live += __popc(mask);
                }
                if (lane == 0) s_n_live[group] = s_n_inframe[group] = n_live;
            }
        } else {
            if (gwarp == 0) {
                const double pref = 1.0 / (2.0 * 3.141592653589793 * p.sigma * p.sigma);
                const double ef = -1.0 / (2.0 * p.sigma * p.sigma);
                int n_live = 0, n_in = 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    bool live = false, inframe = false;
                    double px = 0, py = 0, I = 0, rad = 0;
                    if (j < n) {
                        project_spot(p, sxyz[3 * j], sxyz[3 * j + 1], px, py);
                        I = sint[j];
                        // get_diffraction_pattern keeps the in-frame spots only (simulation2d.py:422-430); the bare
                        // rasteriser also spreads spots lying outside the frame into it
                        if (p.keep_outside || (px >= 0.0 && px < (double)p.W && py >= 0.0 && py < (double)p.H)) {
                            inframe = true;
                            rad = sqrt(log(p.clip / (pref * I)) / ef);  // detector_functions.py:339
                            live = !isnan(rad);
                        }
                    }
                    n_in += __popc(__ballot_sync(0xffffffffu, inframe));
                    const unsigned mask = __ballot_sync(0xffffffffu, live);
                    if (live) {
                        const int d = n_live + __popc(mask & ((1u << lane) - 1u));
                        ss.fx[d] = (float)px;
                        ss.fy[d] = (float)py;
                        ss.amp[d] = (float)(I * pref);
                        // slices :343-352: [max(0, ceil(c - r)), min(n, floor(c + r + 1)))
                        int x_stop = (int)fmax(fmin(floor(px + rad + 1.0), 32000.0), -32000.0);
                        int y_stop = (int)fmax(fmin(floor(py + rad + 1.0), 32000.0), -32000.0);
                        // a negative slice stop counts from the end (numpy / numba): only reachable for spots
                        // outside the frame, i.e. through the bare function
                        if (x_stop < 0) x_stop = max(x_stop + p.W, 0);
                        if (y_stop < 0) y_stop = max(y_stop + p.H, 0);
                        ss.xlo[d] = (short)max(0, (int)fmax(fmin(ceil(px - rad), 32000.0), -32000.0));
                        ss.xhi[d] = (short)(min(p.W, x_stop) - 1);
                        ss.ylo[d] = (short)max(0, (int)fmax(fmin(ceil(py - rad), 32000.0), -32000.0));
                        ss.yhi[d] = (short)(min(p.H, y_stop) - 1);
                    }
                    n_live += __popc(mask);
                }
                if (lane == 0) {
                    s_n_live[group] = n_live;
                    s_n_inframe[group] = n_in;
                }
            }
        }
        group_sync<G>(group);
        const int n_live = s_n_live[group];
        float *img = p.images + (size_t)t * p.H * p.W;

        // reference: no spot in frame -> zeros, returned before the normalisation (simulation2d.py:434-435)
        // (the test is on the IN-FRAME spots; slow-path spots skipped for a NaN radius still count, so an
        // all-skipped pattern is 0 / 0 = NaN in the reference and here)
        const int n_pass = (p.normalize && s_n_inframe[group] > 0) ? 2 : 1;

        if (n_pass == 2) {
            // ---- which regions can hold the maximum?  A region whose upper bound sum_s a_s max(Wy) max(Wx)
            // is below a lower bound of the maximum (the strongest spot's own centre tap) is skipped by the
            // max pass.  With a direct beam this leaves the few regions around it.
            bool prune = FAST && !WIDE;
            float lower = 0.f;
            if (prune) {
                float amin = INFINITY, amax = -INFINITY;
                for (int j = lane; j < n_live; j += 32) {
                    amin = fminf(amin, spot_amp(fs.spot[j]));
                    amax = fmaxf(amax, spot_amp(fs.spot[j]));
                }
                amin = -warp_max(-amin);
                amax = warp_max(amax);
                const float w0 = tap(lut, p.n4, p.radius, 0);
                lower = amax * w0 * w0;
                prune = amin >= 0.f && lower > 0.f;  // (NaN amplitudes fail both tests)
            }
            // every warp of the group sums the bound over its share of the spots (lane = region), the partial
            // sums meet in shared memory in a fixed order
            for (int reg0 = 0; reg0 < n_regions; reg0 += 32) {
                const int reg = reg0 + lane;
                float ub = 0.f;
                if (prune && reg < n_regions) {
                    const int rx0 = (reg % nrx) * RN_RW, ry0 = (reg / nrx) * RN_RH;
                    const int rx1 = min(rx0 + RN_RW, p.W) - 1, ry1 = min(ry0 + RN_RH, p.H) - 1;
                    for (int j = gwarp; j < n_live; j += G) {
                        const uint2 r = fs.spot[j];
                        ub += spot_amp(r) * folded_bound(lut, p.n4, p.radius, rx0, rx1, spot_ix(r), p.W) *
                              folded_bound(lut, p.n4, p.radius, ry0, ry1, spot_iy(r), p.H);
                    }
                }
                if (reg < n_regions) ubound[gwarp * n_regions + reg] = ub;
            }
            group_sync<G>(group);
            for (int reg = gtid; reg < n_regions; reg += GT) {
                float ub = 0.f;
#pragma unroll
                for (int k = 0; k < G; ++k) ub += ubound[k * n_regions + reg];
                flags[reg] = (!prune || ub * 1.001f >= lower) ? 1 : 0;
            }
            group_sync<G>(group);
        }

        float scale = 1.f, vmax_all = INFINITY;
        for (int pass = 0; pass < n_pass; ++pass) {
            const bool store = (pass == n_pass - 1);
            float vmax = -INFINITY;
            for (int reg = gwarp; reg < n_regions; reg += G) {
                if (!store && !flags[reg]) continue;
                const int rx0 = (reg % nrx) * RN_RW, ry0 = (reg / nrx) * RN_RH;
                float acc[8][8];
                bool any, mma = false;
                if (FAST)
                    any = accumulate_region<WIDE, DENSE>(p, fs, n_live, rx0, ry0, lane, acc, hits_s, mma);
                else
                    any = accumulate_slow(p, ss, n_live, rx0, ry0, lane, acc);
                if (!store) {
                    if (!any) {  // no spot reaches the region: all its pixels are exactly 0
                        vmax = fmaxf(vmax, 0.f);
                        continue;
                    }
                    vmax = fmaxf(vmax, region_max<VEC>(p, rx0, ry0, lane, acc, mma));
                } else {
                    float sc = any ? scale : 0.f;
                    // numpy divides: the maximum pixel is exactly 1.  acc * (1 / max) can be 1 ulp off, so the
                    // (few) regions that can hold the maximum scale here and pin that pixel.
                    if (n_pass == 2 && any && flags[reg]) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = (acc[i][j] == vmax_all) ? 1.0f : acc[i][j] * sc;
                        sc = 1.0f;
                    }
                    store_region<VEC>(p, img, rx0, ry0, lane, acc, any, sc, mma);
                }
            }
            if (!store) {
                vmax = warp_max(vmax);
                if (G > 1) {
                    if (lane == 0) s_wmax[warp] = vmax;
                    group_sync<G>(group);
#pragma unroll
                    for (int k = 0; k < G; ++k) vmax = fmaxf(vmax, s_wmax[group * G + k]);
                }
                scale = 1.f / vmax;  // np.divide(pattern, np.max(pattern)), simulation2d.py:440-441
                vmax_all = vmax;
            }
        }
        group_sync<G>(group);  // the group's spot arrays are reused by its next template
        t = t_next;
    }
    // the last CTA to run out of tickets re-arms the counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&p.ticket[1], 1) == (int)gridDim.x - 1) {
        p.ticket[0] = 0;
        p.ticket[1] = 0;
    }
}

template <bool FAST, int G, bool WIDE, bool VEC, bool DENSE = false>
static int launch_render(const RenderParams &p, int group_bytes, size_t lut_bytes, cudaStream_t st) {
    constexpr int NGROUPS = RN_WARPS / G;
    const size_t smem = lut_bytes + (size_t)NGROUPS * group_bytes;
    DS_REQUIRE(smem <= 200 * 1024, "ds_render: cap=%d / radius=%d need %zu bytes of shared memory", p.cap, p.radius,
               smem);
    // (set on every launch: the attribute is per device, and a cached flag would be a data race between streams)
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(render_kernel<FAST, G, WIDE, VEC, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<FAST, G, WIDE, VEC, DENSE>, RN_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    const int want = (p.n_tmpl + NGROUPS - 1) / NGROUPS;
    const int grid = want < num_sms() * per_sm ? want : num_sms() * per_sm;
    render_kernel<FAST, G, WIDE, VEC, DENSE><<<grid, RN_THREADS, smem, st>>>(p, group_bytes);
    return check_launch("ds_render");
}

}  // namespace ds

extern "C" int ds_render_launch_count(int32_t cap, int32_t H, int32_t W, int32_t radius, int32_t fast, double mean_spots_hint) {
    using namespace ds;
    const bool wide = fast == 1 && (radius >= W || radius >= H);
    const int g_opt = option(OPT_RENDER_GROUP);
    const bool g_forced = g_opt == 1 || g_opt == 2 || g_opt == 4 || g_opt == 8;
    if (fast == 1 && !wide && !g_forced && wants_umma(cap, mean_spots_hint) && umma_eligible(H, W, cap, radius)) return 2;
    return 1;
}

extern "C" int64_t ds_render_scratch_bytes(int32_t n_tmpl, int32_t cap) {
    if (n_tmpl < 0 || cap < 0) return -1;
    const int a = ds::umma_record_bytes(cap), b = ds::rows_record_bytes(cap);
    return 16 + (int64_t)n_tmpl * (a > b ? a : b);
}

extern "C" int ds_render(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count, const double *xyz,
                         const double *intensity, int32_t H, int32_t W, double calibration, double cx, double cy,
                         double in_plane_angle_deg, int32_t mirrored, int32_t fast, double sigma, int32_t radius,
                         double clip_threshold, int32_t normalize, float *images, void *scratch,
                         double mean_spots_hint) {
    using namespace ds;
    DS_REQUIRE(n_tmpl >= 0 && cap > 0 && H > 0 && W > 0, "ds_render: bad sizes");
    DS_REQUIRE(H < 16384 && W < 16384, "ds_render: image larger than 16383 px per side");
    DS_REQUIRE(calibration != 0.0, "ds_render: calibration cannot be zero");
    DS_REQUIRE(sigma > 0.0, "ds_render: sigma must be positive");
    DS_REQUIRE((reinterpret_cast<uintptr_t>(images) & 15) == 0, "ds_render: images must be 16-byte aligned");
    DS_REQUIRE(scratch != nullptr || n_tmpl == 0, "ds_render: the scratch buffer is required");
    DS_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "ds_render: scratch must be 16-byte aligned");
    int32_t *ticket = static_cast<int32_t *>(scratch);  // [0..1]: dynamic template hand-out; records from byte 16 on
    if (n_tmpl == 0) return 0;
    RenderParams p;
    p.n_tmpl = n_tmpl;
    p.cap = cap;
    p.H = H;
    p.W = W;
    p.count = count;
    p.xyz = xyz;
    p.intensity = intensity;
    p.cal = calibration;
    p.cx = cx;
    p.cy = cy;
    const double ang = in_plane_angle_deg * (3.141592653589793 / 180.0);
    p.ca = (in_plane_angle_deg == 0.0) ? 1.0 : cos(ang);
    p.sa = (in_plane_angle_deg == 0.0) ? 0.0 : sin(ang);
    p.mirror = mirrored ? -1.0 : 1.0;
    p.sigma = sigma;
    p.clip = clip_threshold;
    p.radius = radius;
    p.normalize = normalize;
    p.images = images;
    p.ticket = ticket;
    p.mean_spots_hint = mean_spots_hint;
    p.table_size = 0;
    p.n4 = 0;
    p.hits_bytes = 0;
    p.mma_min = MMA_MIN_HITS;
    p.mma_tmpl_min = 1 << 30;
    DS_REQUIRE(fast >= 0 && fast <= 2, "ds_render: fast must be 0, 1 or 2");
    p.keep_outside = (fast == 2);
    if (fast == 2) fast = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t lut_bytes = 0;
    int group_bytes;
    const int n_regions = ((W + RN_RW - 1) / RN_RW) * ((H + RN_RH - 1) / RN_RH);
    const bool wide = fast && (radius >= W || radius >= H);
    if (fast) {
        DS_REQUIRE(radius >= 0 && radius <= 2048, "ds_render: radius out of range");
        int ts = 64;
        while (ts < 2 * cap) ts <<= 1;
        p.table_size = ts;
        p.n4 = lut_entries(radius);
        // Tensor-core path for dense regions (DS_RENDER_MMA=0 switches it off).  A region is reached by about `frac`
        // of a template's spots; templates expected to put fewer than mma_min spots into a region skip the hit
        // lists, and launches whose capacity says that few templates are that dense do not carry them at all: the
        // path pays from roughly 100 spots per template on (measured at sigma = 10, 256 x 256: capacity 288 /
        // 177 spots on average 1.75 x faster, capacity 864 / 680 spots 3.7 x; capacity 96 / 39 spots 1.1 - 2 x SLOWER
        // because of the list building in front of mostly sparse regions).
        p.hits_bytes = 0;
        p.mma_min = MMA_MIN_HITS;
        if (option(OPT_RENDER_MMA_MIN) > 1) p.mma_min = option(OPT_RENDER_MMA_MIN);
        const double frac = fmin(1.0, (2.0 * radius + 1 + RN_RW) * (2.0 * radius + 1 + RN_RH) / ((double)W * H));
        p.mma_tmpl_min = (int)ceil(p.mma_min / frac);
        if (option(OPT_RENDER_MMA_TMPL_MIN) >= 0) p.mma_tmpl_min = option(OPT_RENDER_MMA_TMPL_MIN);
        const int force = option(OPT_RENDER_MMA);
        const bool want = force >= 0 ? force != 0 : cap >= 3 * p.mma_tmpl_min;
        if (want && !wide && cap >= p.mma_min && cap <= 4096)
            p.hits_bytes = RN_WARPS * ((2 * cap * 2 + 15) & ~15);  // 2 cap entries: spots + their reflect images
        lut_bytes = lut_smem_bytes(p.n4) + (size_t)p.hits_bytes;
        group_bytes = (int)((size_t)ts * 8 + (size_t)cap * 16);
    } else {
        group_bytes = (int)slow_smem_bytes(cap);
    }
    // stage the spot rows through shared memory when they are small and 16-byte granular
    p.stage = (cap <= 512 && (cap & 1) == 0 && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(intensity) & 15) == 0)
                  ? 1
                  : 0;
    if (option(OPT_RENDER_NOSTAGE) > 0) p.stage = 0;
    if (p.stage) group_bytes += 2 * cap * 32;
    const int group_fixed = (group_bytes + 15) & ~15;  // + flags and per-warp bound partials, which depend on G
    // warps per template, measured on B200 (tools/bench_configs.py): sparse patterns (<= 32 reflections) are
    // write-bound and like the whole CTA on one template (one 256 KB window per CTA); denser ones are
    // phase/latency-bound and like more templates in flight per SM.  DS_RENDER_GROUP overrides.
    int G = cap <= 32 ? 8 : (cap <= 64 ? 4 : 2);
    const int g_opt = option(OPT_RENDER_GROUP);
    const bool g_forced = g_opt == 1 || g_opt == 2 || g_opt == 4 || g_opt == 8;
    if (g_forced) G = g_opt;
    // The ticket words start every launch at zero (a launch that died mid-way cannot poison the next one).
    cudaMemsetAsync(ticket, 0, 2 * sizeof(int32_t), st);
    // Templates of up to 256 x 256 px with more than a handful of reflections: the tcgen05 kernel (render_umma.cu);
    // option render_umma = 0 / 1 forces it off / on.  Measured cross-over on B200 (profiles/r02_k3_variants*.txt): the
    // float32 pipelined kernel wins below ~16 reflections per template (sigma 10), the tcgen05 kernel above.
    if (fast && !wide && !g_forced) {
        if (wants_umma(cap, mean_spots_hint)) {
            // Very dense templates take the row-binned banded product (render_rows.cu), whose cost barely depends on the
            // number of reflections; measured cross-over against the per-reflection product at ~320 reflections per
            // template (sigma 10; profiles/r02_k3_variants_v8.txt).  render_rows = 0 / 1 forces it off / on.
            const int ro = option(OPT_RENDER_ROWS);
            if (ro >= 0 ? ro != 0 : (mean_spots_hint > 0.0 ? mean_spots_hint >= 320.0 : cap >= 640)) {
                const int rc = launch_render_rows(p, static_cast<unsigned char *>(scratch) + 16, st);
                if (rc != 0) return rc < 0 ? rc : 0;
            }
            const int rc = launch_render_umma(p, static_cast<unsigned char *>(scratch) + 16, st);
            if (rc != 0) return rc < 0 ? rc : 0;
        }
    }
    // sparse patterns take the warp-specialised pipelined kernel (render_pipe.cu); render_pipe = 0 disables
    if (fast && !wide && option(OPT_RENDER_PIPE) != 0 && !g_forced) {
        const int rc = launch_render_pipelined(p, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    auto bytes_for = [&](int g) { return (group_fixed + ((n_regions + 3) & ~3) + g * n_regions * 4 + 15) & ~15; };
    while (G < 8 && lut_bytes + (size_t)(RN_WARPS / G) * bytes_for(G) > 96 * 1024) G <<= 1;
    if (wide || (W & 3) != 0) G = 8;
    if (!fast && G != 1) G = 8;  // the sub-pixel path is instantiated for G = 1 and G = 8 only
    group_bytes = bytes_for(G);
#define DS_RN(F, GG, WD, V) launch_render<F, GG, WD, V>(p, group_bytes, lut_bytes, st)
#define DS_RND(F, GG, WD, V) launch_render<F, GG, WD, V, true>(p, group_bytes, lut_bytes, st)
    // the tensor-core path exists in the aligned, non-wide fast instantiations with G >= 2
    const bool dense = p.hits_bytes > 0;
    // rows that are not 16-byte aligned (W % 4 != 0) and kernels wider than the image take the
    // general-purpose instantiations; the common case gets the lean ones
    if ((W & 3) != 0) return fast ? (wide ? DS_RN(true, 8, true, false) : DS_RN(true, 8, false, false))
                                  : DS_RN(false, 8, false, false);
    if (wide) return DS_RN(true, 8, true, true);
    if (fast) {
        switch (G) {
            case 1: return DS_RN(true, 1, false, true);
            case 2: return dense ? DS_RND(true, 2, false, true) : DS_RN(true, 2, false, true);
            case 4: return dense ? DS_RND(true, 4, false, true) : DS_RN(true, 4, false, true);
            default: return dense ? DS_RND(true, 8, false, true) : DS_RN(true, 8, false, true);
        }
    } else {
        switch (G) {
            case 1: return DS_RN(false, 1, false, true);
            default: return DS_RN(false, 8, false, true);
        }
    }
#undef DS_RN
#undef DS_RND
}
