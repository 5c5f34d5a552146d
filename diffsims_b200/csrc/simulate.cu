// K2 -- fused rotate / excitation error / shape factor / cull / threshold over (rotation x g).
//
// Replaces the body of the reference's rotation loop (diffsims/generators/simulation_generator.py:211-241):
//   rotate_with_basis           crystallography/_diffracting_vector.py:127-161   g_lab = R g
//   get_intersecting_reflections simulation_generator.py:319-412                  s, |s| < s_max, shape factor
//   prefactor * |F|^2            utils/sim_utils.py:353                           I = shape(s) * I0[g]
//   minimum_intensity threshold  simulation_generator.py:237                      I > max(I) * min_intensity
// and the same arithmetic of the old API (generators/diffraction_generator.py:247-324).
//
// Mapping: one warp per rotation, persistent CTAs of 8 warps.  The per-phase g table (float4 rows
// gx, gy, gz, |g|^2) is staged into shared memory with cp.async.bulk (TMA engine) -- once per CTA when it
// fits, double-buffered tiles otherwise.  Each lane tests four g per step in float32 with a safety margin
// (only z' = R[2,:].g is needed because rotation preserves |g|; the Ewald test |s| < t is evaluated in the
// cancellation-free, sqrt-free form f(z'+t) < 0 < f(z'-t), f(u) = r^2 + u (u - 2 r_s)); a warp ballot +
// prefix sum compacts the
// candidates into a per-warp list, and full warps of candidates are then refined in float64 (full rotation,
// the reference's own excitation-error expression, strict cut, shape factor).  float64 is required for the
// refine: the Lorentzian's sensitivity dI/I ~ 180 ds near s_max needs ds < 5e-8.
// Survivors are appended to the padded output row in g-table order; a second in-place ballot compaction
// applies the max-relative intensity threshold.
#include <stdlib.h>

#include "common.cuh"

namespace ds {

constexpr int SIM_WARPS = 8;
constexpr int SIM_THREADS = SIM_WARPS * 32;
constexpr int SIM_RESIDENT_MAX_G = 6144;  // 96 KB of float4 rows
constexpr int SIM_TILE_G = 3072;          // streaming: 2 x 48 KB

struct SimParams {
    int n_rot, n_g, cap;
    const double *quat;
    const double *g_xyz;
    const float4 *g_f32;
    const double *g_I0;
    double rs;  // 1 / wavelength
    double s_max, width, minima, prec, min_intensity;
    float coarse_margin;
    int model;
    int n_quad;  // > 0: average `model` over the precession circle with n_quad midpoint nodes on [0, pi]
    // scan-line mode (n_lines > 0): the table is a set of lattice lines g0 + i * step, i = 0 .. count-1
    int n_lines;
    const float4 *line_g0;   // [n_lines] (g0.x, g0.y, g0.z, unused)
    const int *line_start;   // [n_lines + 1] first table index of each line
    float step[3];
    float z_lo, z_hi;        // slab of z' that can hold a reflection (cut + margin included)
    int *count;
    int *g_index;
    double *xyz;
    double *intensity;
    double *exc;
    int *max_count;
};

__device__ __forceinline__ double shape_factor(int model, double s, double w, double minima, double r_spot,
                                               double prec) {
    const double PI = 3.141592653589793;
    switch (model) {
        case DS_SHAPE_LINEAR: {  // shape_factor_models.py:52-73
            const double sf = 1.0 - fabs(s) / w;
            return sf < 0.0 ? 0.0 : sf;
        }
        case DS_SHAPE_SINC:
        case DS_SHAPE_SIN2C: {  // :76-123 (where=denom != 0 leaves 0 at s == 0)
            const double fac = PI * minima / w;
            const double den = fac * s;
            const double v = (den != 0.0) ? fabs(sin(fac * s) / den) : 0.0;
            return model == DS_SHAPE_SINC ? v : v * v;
        }
        case DS_SHAPE_ATANC: {  // :126-151 (nan_to_num(nan=1) at s == 0)
            const double fac = PI * minima / fabs(w);
            const double x = fac * s;
            return (x != 0.0) ? atan(x) / x : 1.0;
        }
        case DS_SHAPE_LORENTZIAN: {  // :154-180
            const double sigma = PI / w;
            return sigma / (PI * (sigma * sigma * (s * s) + 1.0)) * w;
        }
        case DS_SHAPE_LORENTZIAN_PRECESSION: {  // :183-219
            const double sigma = PI / w;
            const double u = sigma * sigma * (r_spot * r_spot * (prec * prec) - s * s) + 1.0;
            const double z = sqrt(u * u + 4.0 * (sigma * sigma) * (s * s));
            return (sigma / PI) * sqrt(2.0 * (u + z) / (z * z));
        }
        default:  // binary, return-s
            return 1.0;
    }
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    // volatile: the tile buffers are refilled by the async proxy between tiles
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// float32 coarse test of one g (z' = R[2,:].g, r^2 = |g|^2 - z'^2): cancellation-free and sqrt-free,
//   |s| < t  <=>  f(z'+t) < 0 < f(z'-t),  f(u) = r^2 + u (u - 2 r_s).
// Both f values are linear in z': f(z' -+ t) = q +- d with q = |g|^2 + t^2 - 2 r_s z' and d = 2 t (r_s - z') > 0, so the
// test is |q| < d: two FFMA, one FADD and one compare after the three FFMA of z'.  Rows with |g|^2 = +inf (extinct rows,
// tail padding) give q = +inf and fail.  With precession the two-surface test of simulation_generator.py:365-375 in the
// same algebra around the tilted sphere centre (P_t, P_z).
struct CoarseConst {
    float two_rs, thr, t2, two_t, two_t_rs, P_z, P_t;
    bool prec_on;
};
__device__ __forceinline__ bool coarse_test(float z, float g2, const CoarseConst &c) {
    if (!c.prec_on) {
        const float q = fmaf(-c.two_rs, z, g2 + c.t2), d = fmaf(-c.two_t, z, c.two_t_rs);
        return fabsf(q) < d;
    }
    const float r2 = fmaf(-z, z, g2);
    const float u = z + c.thr, v = z - c.thr;
    const float r = sqrtf(fmaxf(r2, 0.0f)), two_rpt = 2.0f * r * c.P_t;
    return (r2 + two_rpt + v * (v - 2.0f * c.P_z) >= 0.0f) && (r2 - two_rpt + u * (u - 2.0f * c.P_z) <= 0.0f);
}
__device__ __forceinline__ CoarseConst coarse_const(double rs, double s_max, float margin, double prec, bool general) {
    CoarseConst c;
    c.two_rs = 2.0f * (float)rs;
    c.thr = (float)s_max + margin;
    c.t2 = c.thr * c.thr;
    c.two_t = 2.0f * c.thr;
    c.two_t_rs = c.two_t * (float)rs;
    c.prec_on = general && prec != 0.0;
    c.P_z = general ? (float)(rs * cos(prec)) : 0.f;
    c.P_t = general ? (float)(rs * sin(prec)) : 0.f;
    return c;
}

// scan-line mode rebuilds g from the line tables; rows marked extinct by ds_pack_gtable (|g|^2 = +inf in the packed
// table) are dropped among the few rows that pass the coarse test
__device__ __forceinline__ bool live_row(const SimParams &p, int index) {
    return __ldg(reinterpret_cast<const float *>(p.g_f32) + 4 * (size_t)index + 3) < INFINITY;
}

struct WarpState {
    // float64 active rotation matrix, row-major
    double m[9];
    int n_out;     // reflections that passed the excitation-error test so far
    double max_I;  // running max of their intensities
};

// Float64 evaluation of up to 32 candidates (one per lane): full rotation, the reference's own excitation-error
// expression, the strict cut, shape factor and intensity.  MODEL >= 0 fixes the shape factor at compile time (no
// precession); MODEL < 0 is the general kernel (runtime model, precession cut, closed-form or numerically averaged
// precession shape factor -- the average is warp-cooperative, so all 32 lanes must call this).  WITH_I = false stops
// after the cut (x, y, z, s only).
struct Refined {
    double x, y, z, s, I;
    bool keep;
};
template <int MODEL, bool WITH_I>
__device__ __forceinline__ Refined refine_eval(const SimParams &p, const double (&m)[9], bool have, int gi, int lane,
                                               const double *__restrict__ s_cos) {
    constexpr bool GENERAL = MODEL < 0;
    const int model = GENERAL ? p.model : MODEL;
    Refined r;
    r.keep = false;
    r.x = r.y = r.z = r.s = r.I = 0.0;
    double r_spot = 0;
    if (have) {
        const double gx = __ldg(p.g_xyz + 3 * (size_t)gi), gy = __ldg(p.g_xyz + 3 * (size_t)gi + 1),
                     gz = __ldg(p.g_xyz + 3 * (size_t)gi + 2);
        r.x = m[0] * gx + m[1] * gy + m[2] * gz;
        r.y = m[3] * gx + m[4] * gy + m[5] * gz;
        r.z = m[6] * gx + m[7] * gy + m[8] * gz;
        // simulation_generator.py:355-360, evaluated as the reference writes it
        r_spot = sqrt(r.x * r.x + r.y * r.y);
        const double z_sphere = -sqrt(p.rs * p.rs - r_spot * r_spot) + p.rs;
        r.s = z_sphere - r.z;
        if (!GENERAL || p.prec == 0.0) {
            r.keep = fabs(r.s) < p.s_max;  // :364 strict
        } else {                           // :365-375
            const double P_z = p.rs * cos(p.prec), P_t = p.rs * sin(p.prec);
            const double up = P_z - sqrt(p.rs * p.rs - (r_spot + P_t) * (r_spot + P_t));
            const double dn = P_z - sqrt(p.rs * p.rs - (r_spot - P_t) * (r_spot - P_t));
            r.keep = (r.z - p.s_max <= up) && (r.z + p.s_max >= dn);
        }
    }
    if (!WITH_I) return r;
    double sf = 1.0;
    if (GENERAL && p.n_quad > 0) {
        // _shape_factor_precession (shape_factor_models.py:222-269): (1 / 2 pi) int_0^2pi f(s + r phi cos t) dt.
        // The integrand is even and periodic in t, so the midpoint rule on [0, pi] (Gauss-Chebyshev in
        // u = cos t) converges geometrically for the smooth models and as 1/n^2 for the kinked ones; the
        // whole warp integrates one candidate at a time over a cosine table in shared memory.
        for (int c = 0; c < 32; ++c) {
            if (!__shfl_sync(0xffffffffu, (int)r.keep, c)) continue;
            const double sc = __shfl_sync(0xffffffffu, r.s, c);
            const double amp = __shfl_sync(0xffffffffu, r_spot, c) * p.prec;
            double acc = 0.0;
            for (int j = lane; j < p.n_quad; j += 32)
                acc += shape_factor(model, sc + amp * s_cos[j], p.width, p.minima, 0.0, 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == c) sf = acc / (double)p.n_quad;
        }
    } else if (r.keep) {
        sf = shape_factor(model, r.s, p.width, p.minima, r_spot, GENERAL ? p.prec : 0.0);
    }
    if (r.keep) r.I = sf * __ldg(p.g_I0 + gi);
    return r;
}

// Refine up to 32 candidates and append the survivors to the output row.
template <int MODEL>
__device__ __forceinline__ void refine(const SimParams &p, WarpState &w, int rot, bool have, int gi, int lane,
                                       const double *__restrict__ s_cos) {
    const Refined r = refine_eval<MODEL, true>(p, w.m, have, gi, lane, s_cos);
    const unsigned mask = __ballot_sync(0xffffffffu, r.keep);
    const int slot = w.n_out + __popc(mask & ((1u << lane) - 1u));
    if (r.keep && slot < p.cap) {
        const size_t o = (size_t)rot * p.cap + slot;
        p.xyz[3 * o + 0] = r.x;
        p.xyz[3 * o + 1] = r.y;
        p.xyz[3 * o + 2] = r.z;
        p.intensity[o] = r.I;
        p.g_index[o] = gi;
        if (p.exc) p.exc[o] = r.s;
    }
    w.n_out += __popc(mask);
    w.max_I = fmax(w.max_I, warp_max(r.keep ? r.I : -INFINITY));
}

// Active rotation matrix of a unit quaternion (a, b, c, d), row-major float64.
__device__ __forceinline__ void quat_matrix(const double *__restrict__ q, double (&m)[9]) {
    const double a = q[0], b = q[1], c = q[2], d = q[3];
    m[0] = a * a + b * b - c * c - d * d;
    m[1] = 2 * (b * c - a * d);
    m[2] = 2 * (b * d + a * c);
    m[3] = 2 * (b * c + a * d);
    m[4] = a * a - b * b + c * c - d * d;
    m[5] = 2 * (c * d - a * b);
    m[6] = 2 * (b * d - a * c);
    m[7] = 2 * (c * d + a * b);
    m[8] = a * a - b * b - c * c + d * d;
}

// The minimum-intensity cut of one stored row, in place and in order (simulation_generator.py:237): one warp.
template <int MODEL>
__device__ __forceinline__ int threshold_row(const SimParams &p, int rot, int n_out, double max_I, int lane) {
    constexpr bool GENERAL = MODEL < 0;
    const int n_stored = min(n_out, p.cap);
    if ((GENERAL ? p.model : MODEL) == DS_SHAPE_NONE_RETURN_S || p.min_intensity < 0.0) return n_stored;  // cut disabled
    int n_keep = 0;
    const double cut = max_I * p.min_intensity;
    const size_t row = (size_t)rot * p.cap;
    for (int j0 = 0; j0 < n_stored; j0 += 32) {
        const int j = j0 + lane;
        bool keep = false;
        double x = 0, y = 0, z = 0, I = 0, s = 0;
        int gi = 0;
        if (j < n_stored) {
            I = p.intensity[row + j];
            keep = I > cut;
            if (keep) {
                x = p.xyz[3 * (row + j)];
                y = p.xyz[3 * (row + j) + 1];
                z = p.xyz[3 * (row + j) + 2];
                gi = p.g_index[row + j];
                if (p.exc) s = p.exc[row + j];
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int dst = n_keep + __popc(mask & ((1u << lane) - 1u));
        __syncwarp();
        if (keep && dst != j) {
            p.xyz[3 * (row + dst)] = x;
            p.xyz[3 * (row + dst) + 1] = y;
            p.xyz[3 * (row + dst) + 2] = z;
            p.intensity[row + dst] = I;
            p.g_index[row + dst] = gi;
            if (p.exc) p.exc[row + dst] = s;
        }
        n_keep += __popc(mask);
        __syncwarp();
    }
    return n_keep;
}

template <int MODEL, bool LINES>
__global__ void __launch_bounds__(SIM_THREADS, MODEL < 0 ? 2 : 3) simulate_kernel(const SimParams p, const int n_tiles,
                                                                  const int tile_g, const int tile_alloc) {
    constexpr bool GENERAL = MODEL < 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // `tile_alloc` rows per buffer: tile_g rounded up to a multiple of 128 (the brute-force scan reads 128 rows per step
    // without bounds checks; the rows past a tile's end hold |g|^2 = +inf and fail the test).  In scan-line mode the host
    // passes tile_g = tile_alloc = bytes of the line tables / 16, n_tiles = 1.
    float4 *s_tile[2] = {reinterpret_cast<float4 *>(smem_raw),
                         reinterpret_cast<float4 *>(smem_raw) + (n_tiles > 1 ? tile_alloc : 0)};
    double *s_cos = reinterpret_cast<double *>(smem_raw + (size_t)tile_alloc * 16 * (n_tiles > 1 ? 2 : 1));
    if (GENERAL)
        for (int j = threadIdx.x; j < p.n_quad; j += blockDim.x) s_cos[j] = cospi(((double)j + 0.5) / (double)p.n_quad);
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_list[SIM_WARPS][64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t phase[2] = {0, 0};
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int buf) {
        if (threadIdx.x == 0) {
            const int n = min(tile_g, p.n_g - tile * tile_g);
            mbar_expect_tx(&s_bar[buf], (uint32_t)n * 16u);
            bulk_g2s(s_tile[buf], p.g_f32 + (size_t)tile * tile_g, (uint32_t)n * 16u, &s_bar[buf]);
        }
    };

    const int *s_lstart = nullptr;
    if (LINES) {
        // line table (g0 as float4) + line starts, resident for the CTA's lifetime
        const uint32_t b0 = (uint32_t)p.n_lines * 16u, b1 = (uint32_t)((p.n_lines + 1 + 3) & ~3) * 4u;
        int *dst1 = reinterpret_cast<int *>(smem_raw + b0);
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar[0], b0 + b1);
            bulk_g2s(smem_raw, p.line_g0, b0, &s_bar[0]);
            bulk_g2s(dst1, p.line_start, b1, &s_bar[0]);
        }
        mbar_wait(&s_bar[0], 0);
        s_lstart = dst1;
    } else if (n_tiles == 1) {  // resident table: one bulk copy for the CTA's lifetime
        issue(0, 0);
        for (int i = p.n_g + (int)threadIdx.x; i < tile_alloc; i += blockDim.x) s_tile[0][i] = make_float4(0.f, 0.f, 0.f, INFINITY);
        mbar_wait(&s_bar[0], 0);
        __syncthreads();
    }
    // streaming: the last tile is short; its buffer's rows up to the next multiple of 128 are re-padded every time it is
    // issued (the full tile that used the buffer before overwrote them)
    const int n_last = p.n_g - (n_tiles - 1) * tile_g;
    auto pad_last = [&](int buf) {
        for (int i = n_last + (int)threadIdx.x; i < ((n_last + 127) & ~127); i += blockDim.x)
            s_tile[buf][i] = make_float4(0.f, 0.f, 0.f, INFINITY);
    };

    const CoarseConst cc = coarse_const(p.rs, p.s_max, p.coarse_margin, p.prec, GENERAL);
    const uint32_t tile_base_s = smem_u32(smem_raw);
    int local_max_count = 0;

    // rotations per CTA = warps per CTA: 8 normally, fewer when the launch has too few rotations to fill the SMs
    const int wpb = blockDim.x >> 5;
    const int n_batches = (p.n_rot + wpb - 1) / wpb;
    for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        const int rot = batch * wpb + warp;
        const bool active = rot < p.n_rot;
        WarpState w;
        w.n_out = 0;
        w.max_I = -INFINITY;
        float mz0 = 0, mz1 = 0, mz2 = 0;
        if (active) {
            quat_matrix(p.quat + 4 * (size_t)rot, w.m);
            mz0 = (float)w.m[6];
            mz1 = (float)w.m[7];
            mz2 = (float)w.m[8];
        }
        int n_list = 0;
        int *list = s_list[warp];
        const uint32_t list_s = smem_u32(list);

        // append `cand` lanes' table indices to the warp's candidate list in lane order; refine full warps
        auto append = [&](bool cand, int index) {
            const unsigned mask = __ballot_sync(0xffffffffu, cand);
            if (mask == 0u) return;
            if (cand) sts_u32(list_s + 4u * (uint32_t)(n_list + __popc(mask & ((1u << lane) - 1u))), (uint32_t)index);
            n_list += __popc(mask);
            __syncwarp();
            if (n_list >= 32) {
                refine<MODEL>(p, w, rot, true, list[lane], lane, s_cos);
                const int rest = n_list - 32;
                const int carry = (lane < rest) ? list[32 + lane] : 0;
                __syncwarp();
                list[lane] = carry;
                n_list = rest;
                __syncwarp();
            }
        };

        if (LINES) {
            // ---- scan-line cull: a lattice line g0 + i step crosses the slab z_lo <= z' <= z_hi in one short
            // index interval, found in closed form; only those entries see the exact float32 test.  One lane per
            // line, 32 lines per step, candidates emitted in table order (lane order, ascending i).
            if (active) {
                const float bz = fmaf(mz0, p.step[0], fmaf(mz1, p.step[1], mz2 * p.step[2]));  // dz' per index
                const bool parallel = fabsf(bz) < 1e-3f;
                const float inv_b = parallel ? 0.0f : 1.0f / bz;
                for (int L0 = 0; L0 < p.n_lines; L0 += 32) {
                    const int L = L0 + lane;
                    int start = 0, cnt = 0, ilo = 0;
                    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (L < p.n_lines) {
                        g0 = lds_f4(tile_base_s + 16u * (uint32_t)L);
                        start = s_lstart[L];
                        const int len = s_lstart[L + 1] - start;
                        const float a = fmaf(mz0, g0.x, fmaf(mz1, g0.y, mz2 * g0.z));
                        int ihi;
                        if (!parallel) {
                            float t0 = (p.z_lo - a) * inv_b, t1 = (p.z_hi - a) * inv_b;
                            if (t0 > t1) {
                                const float tmp = t0;
                                t0 = t1;
                                t1 = tmp;
                            }
                            ilo = max(0, (int)ceilf(t0 - 4e-3f));
                            ihi = min(len - 1, (int)floorf(t1 + 4e-3f));
                        } else {  // the line runs (almost) parallel to the slab: all of it or nothing
                            const float slack = (float)len * fabsf(bz);
                            const bool inside = a >= p.z_lo - slack && a <= p.z_hi + slack;
                            ilo = 0;
                            ihi = inside ? len - 1 : -1;
                        }
                        cnt = max(0, ihi - ilo + 1);
                    }
                    const int cmax = warp_max(cnt);
                    if (cmax == 0) continue;
                    if (cmax <= 4) {
                        // common case: every lane tests its (at most four) entries itself, survivors are written
                        // at prefix-sum offsets so that the list stays in table order
                        unsigned bits = 0;
                        for (int r = 0; r < cnt; ++r) {
                            const float fi = (float)(ilo + r);
                            const float gx = fmaf(fi, p.step[0], g0.x), gy = fmaf(fi, p.step[1], g0.y),
                                        gz = fmaf(fi, p.step[2], g0.z);
                            const float z = fmaf(mz0, gx, fmaf(mz1, gy, mz2 * gz));
                            if (coarse_test(z, fmaf(gx, gx, fmaf(gy, gy, gz * gz)), cc) &&
                                live_row(p, start + ilo + r))
                                bits |= 1u << r;
                        }
                        const int mine = __popc(bits);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += up;
                        }
                        const int total = __shfl_sync(0xffffffffu, incl, 31);
                        if (total == 0) continue;
                        if (n_list + total <= 64) {
                            int at = n_list + incl - mine;
                            while (bits) {
                                const int r = __ffs(bits) - 1;
                                bits &= bits - 1;
                                sts_u32(list_s + 4u * (uint32_t)at++, (uint32_t)(start + ilo + r));
                            }
                            n_list += total;
                            __syncwarp();
                            while (n_list >= 32) {
                                refine<MODEL>(p, w, rot, true, list[lane], lane, s_cos);
                                const int rest = n_list - 32;
                                const int carry = (lane < rest) ? list[32 + lane] : 0;
                                __syncwarp();
                                list[lane] = carry;
                                n_list = rest;
                                __syncwarp();
                            }
                            continue;
                        }
                        // (more than the list can take at once: fall through to the line-by-line path)
                    }
                    // long intervals (zone-axis-like orientations) or a burst: the warp scans one line at a time
                    for (int src = 0; src < 32; ++src) {
                        const int c = __shfl_sync(0xffffffffu, cnt, src);
                        if (c == 0) continue;
                        const int first = __shfl_sync(0xffffffffu, start + ilo, src);
                        const int i_first = __shfl_sync(0xffffffffu, ilo, src);
                        const float hx = __shfl_sync(0xffffffffu, g0.x, src), hy = __shfl_sync(0xffffffffu, g0.y, src),
                                    hz = __shfl_sync(0xffffffffu, g0.z, src);
                        for (int m = 0; m < c; m += 32) {
                            const int r = m + lane;
                            bool cand = false;
                            if (r < c) {
                                const float fi = (float)(i_first + r);
                                const float gx = fmaf(fi, p.step[0], hx), gy = fmaf(fi, p.step[1], hy),
                                            gz = fmaf(fi, p.step[2], hz);
                                const float z = fmaf(mz0, gx, fmaf(mz1, gy, mz2 * gz));
                                cand = coarse_test(z, fmaf(gx, gx, fmaf(gy, gy, gz * gz)), cc) &&
                                       live_row(p, first + r);
                            }
                            append(cand, first + r);
                        }
                    }
                }
            }
        } else {
            if (n_tiles > 1) issue(0, 0);
            for (int t = 0; t < n_tiles; ++t) {
                const int buf = (n_tiles > 1) ? (t & 1) : 0;
                if (n_tiles > 1) {
                    if (t + 1 < n_tiles) {
                        issue(t + 1, (t + 1) & 1);
                        if (t + 2 == n_tiles) pad_last((t + 1) & 1);  // (read after the __syncthreads that ends this tile)
                    }
                    mbar_wait(&s_bar[buf], phase[buf]);
                    phase[buf] ^= 1;
                }
                const int n = min(tile_g, p.n_g - t * tile_g);
                const uint32_t tile_s = tile_base_s + (buf ? (uint32_t)tile_alloc * 16u : 0u);
                if (active) {
                    for (int i0 = 0; i0 < n; i0 += 128) {
                        // four independent g per lane: loads and tests overlap, ballots are consumed in table order
                        bool cands[4];
                        float4 gk[4];
    #pragma unroll
                        for (int k = 0; k < 4; ++k)
                            gk[k] = lds_f4(tile_s + 16u * (uint32_t)(i0 + 32 * k + lane));
                        bool any = false;
    #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 g = gk[k];
                            const float z = fmaf(mz0, g.x, fmaf(mz1, g.y, mz2 * g.z));
                            cands[k] = coarse_test(z, g.w, cc);
                            any |= cands[k];
                        }
                        if (!__any_sync(0xffffffffu, any)) continue;
    #pragma unroll
                        for (int k = 0; k < 4; ++k) append(cands[k], t * tile_g + i0 + 32 * k + lane);
                    }
                }
                if (n_tiles > 1) __syncthreads();  // everyone is done with `buf` before it is refilled
            }
        }
        if (active) {
            if (n_list > 0) refine<MODEL>(p, w, rot, lane < n_list, lane < n_list ? list[lane] : 0, lane, s_cos);
            __syncwarp();
            local_max_count = max(local_max_count, w.n_out);
            const int n_keep = threshold_row<MODEL>(p, rot, w.n_out, w.max_I, lane);
            if (lane == 0) p.count[rot] = n_keep;
        }
    }
    local_max_count = warp_max(local_max_count);
    if (lane == 0 && local_max_count > 0) atomicMax(p.max_count, local_max_count);
}

// ---------------------------------------------------------------------------------------------------
// Large tables (>= 4096 rows): one CTA per rotation, the table split across the CTA's warps.
//
// With one warp per rotation a launch of a few hundred rotations over a 10^5-row table is bound by one warp's latency over
// the whole table, and at any rotation count every CTA streams the table through shared memory tile by tile.  Here the eight warps of a CTA scan eight contiguous slices of the table (straight from L2 / L1: the
// CTAs of an SM run in step and share the lines), so the rotation's reflections come out in table order as slice 0,
// slice 1, ...:
//   1. scan: float32 coarse test, candidate row indices into the stash (shared memory).  The stash is one pool of
//      64-entry chunks that the warps draw from as they fill up -- the slab of a zone-axis-like orientation puts most
//      candidates into one or two slices;
//   2. evaluate the candidates in float64 (refine_eval), intensities into the stash (NaN = failed the cut);
//      CTA-wide: number of reflections before the intensity cut (max_count) and their maximum -> the cut;
//   3. count the survivors per warp, CTA prefix sum -> each warp's offset in the output row;
//   4. re-evaluate the survivors' geometry (bit-identical: same code) and store them at their final positions.
// A rotation with more candidates than the pool holds goes to warp 0, which runs the one-warp algorithm of
// simulate_kernel straight from global memory.  Results are identical to simulate_kernel's whenever max_count <= cap (the
// only case callers accept).
// ---------------------------------------------------------------------------------------------------
constexpr int SIM_CHUNK = 64;        // stash entries per chunk
constexpr int SIM_MAX_CHUNKS = 255;  // chunk ids are bytes

template <int MODEL, bool LINES>
__global__ void __launch_bounds__(SIM_THREADS, 4) simulate_cta_kernel(const SimParams p, const int n_chunks) {
    constexpr bool GENERAL = MODEL < 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [n_chunks][64] intensities (double) | same shape candidate rows (int) | [warps][n_chunks] chunk ids | cosine table
    double *s_I = reinterpret_cast<double *>(smem_raw);
    int *s_cand = reinterpret_cast<int *>(s_I + (size_t)n_chunks * SIM_CHUNK);
    unsigned char *s_ids = reinterpret_cast<unsigned char *>(s_cand + (size_t)n_chunks * SIM_CHUNK);
    double *s_cos = reinterpret_cast<double *>(s_ids + (((size_t)SIM_WARPS * n_chunks + 15) & ~(size_t)15));
    if (GENERAL)
        for (int j = threadIdx.x; j < p.n_quad; j += blockDim.x) s_cos[j] = cospi(((double)j + 0.5) / (double)p.n_quad);
    __shared__ int s_ncand[SIM_WARPS], s_nout[SIM_WARPS], s_nkeep[SIM_WARPS];
    __shared__ double s_wmax[SIM_WARPS];
    __shared__ int s_next;      // next free chunk of the pool
    __shared__ int s_list[64];  // fall-back: warp 0's candidate list

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CoarseConst cc = coarse_const(p.rs, p.s_max, p.coarse_margin, p.prec, GENERAL);
    // this warp's slice of the table: a multiple of 128 rows
    const int per_warp = ((p.n_g + SIM_WARPS * 128 - 1) / (SIM_WARPS * 128)) * 128;
    const int row_lo = min(p.n_g, warp * per_warp), row_hi = min(p.n_g, row_lo + per_warp);
    const bool cut_off = (GENERAL ? p.model : MODEL) == DS_SHAPE_NONE_RETURN_S || p.min_intensity < 0.0;
    unsigned char *my_ids = s_ids + (size_t)warp * n_chunks;
    // stash position of this warp's j-th candidate
    auto at = [&](int j) { return (int)my_ids[j / SIM_CHUNK] * SIM_CHUNK + (j % SIM_CHUNK); };
    int local_max_count = 0;

    for (int rot = blockIdx.x; rot < p.n_rot; rot += gridDim.x) {
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();  // (also: the previous rotation's stash has been consumed)
        double m[9];
        quat_matrix(p.quat + 4 * (size_t)rot, m);
        const float mz0 = (float)m[6], mz1 = (float)m[7], mz2 = (float)m[8];
        // ---- 1. scan this warp's slice
        int n_c = 0, n_have = 0;  // candidates so far (counted on even when the pool is exhausted); chunks held
        bool full = false;
        // (the rows of the next step are requested before this step's are tested: a slice streams from L2 at a few hundred
        // nanoseconds per round trip, and a CTA per SM has only eight warps to hide it; slices are whole 128-row steps
        // except at the end of the table, which takes the guarded step below)
        // candidates of one ballot, in lane order, to the end of this warp's stash (one more chunk from the pool when needed)
        auto append = [&](bool cand, int index) {
            const unsigned mask = __ballot_sync(0xffffffffu, cand);
            if (mask == 0u) return;
            const int n_new = __popc(mask);
            if (!full && n_c + n_new > n_have * SIM_CHUNK) {  // (at most one more chunk: 32 <= SIM_CHUNK)
                int id = 0;
                if (lane == 0) id = atomicAdd(&s_next, 1);
                id = __shfl_sync(0xffffffffu, id, 0);
                if (id < n_chunks) {
                    if (lane == 0) my_ids[n_have] = (unsigned char)id;
                    ++n_have;
                    __syncwarp();
                } else {
                    full = true;
                }
            }
            const int j = n_c + __popc(mask & ((1u << lane) - 1u));
            if (cand && j < n_have * SIM_CHUNK) s_cand[at(j)] = index;
            n_c += n_new;
        };
        if (LINES) {
            // ---- scan-line cull (tables made of lattice lines g0 + i step): a line crosses the slab z_lo <= z' <= z_hi in
            // one index interval, found in closed form; only those rows see the float32 test.  Each warp takes a contiguous
            // run of lines, 32 at a time; the intervals of a batch are expanded into (line, i) items in table order and
            // dealt to the lanes 32 items at a time (a line nearly parallel to the slab contributes many, most a few).
            const int lines_per_warp = ((p.n_lines + SIM_WARPS * 32 - 1) / (SIM_WARPS * 32)) * 32;
            const int L_lo = min(p.n_lines, warp * lines_per_warp), L_hi = min(p.n_lines, L_lo + lines_per_warp);
            const float bz = fmaf(mz0, p.step[0], fmaf(mz1, p.step[1], mz2 * p.step[2]));  // dz' per index
            const bool parallel = fabsf(bz) < 1e-3f;
            const float inv_b = parallel ? 0.0f : 1.0f / bz;
            // (the next batch's line records are requested before this batch is expanded)
            float4 g0_n = make_float4(0.f, 0.f, 0.f, 0.f);
            int start_n = 0, end_n = 0;
            if (L_lo + lane < L_hi) {
                g0_n = __ldg(p.line_g0 + L_lo + lane);
                start_n = __ldg(p.line_start + L_lo + lane);
                end_n = __ldg(p.line_start + L_lo + lane + 1);
            }
            for (int L0 = L_lo; L0 < L_hi; L0 += 32) {
                const int Ln = L0 + lane;
                int first = 0, cnt = 0, ilo = 0;
                const float4 g0 = g0_n;
                const int start = start_n, len = end_n - start_n;
                if (Ln + 32 < L_hi) {
                    g0_n = __ldg(p.line_g0 + Ln + 32);
                    start_n = __ldg(p.line_start + Ln + 32);
                    end_n = __ldg(p.line_start + Ln + 33);
                }
                if (Ln < L_hi) {
                    const float a = fmaf(mz0, g0.x, fmaf(mz1, g0.y, mz2 * g0.z));
                    int ihi;
                    if (!parallel) {
                        float t0 = (p.z_lo - a) * inv_b, t1 = (p.z_hi - a) * inv_b;
                        if (t0 > t1) {
                            const float tmp = t0;
                            t0 = t1;
                            t1 = tmp;
                        }
                        ilo = max(0, (int)ceilf(t0 - 4e-3f));
                        ihi = min(len - 1, (int)floorf(t1 + 4e-3f));
                    } else {  // the line runs (almost) parallel to the slab: all of it or nothing
                        const float slack = (float)len * fabsf(bz);
                        ilo = 0;
                        ihi = (a >= p.z_lo - slack && a <= p.z_hi + slack) ? len - 1 : -1;
                    }
                    cnt = max(0, ihi - ilo + 1);
                    first = start + ilo;
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int excl = incl - cnt;
                for (int t0 = 0; t0 < total; t0 += 32) {
                    const int t = t0 + lane;
                    // the line of item t: the first lane whose inclusive count exceeds t
                    int src = 0;
#pragma unroll
                    for (int stp = 16; stp > 0; stp >>= 1)
                        if (__shfl_sync(0xffffffffu, incl, src + stp - 1) <= t) src += stp;
                    src = min(src, 31);
                    const int r = t - __shfl_sync(0xffffffffu, excl, src);
                    const int index = __shfl_sync(0xffffffffu, first, src) + r;
                    const float fi = (float)(__shfl_sync(0xffffffffu, ilo, src) + r);
                    const float hx = __shfl_sync(0xffffffffu, g0.x, src), hy = __shfl_sync(0xffffffffu, g0.y, src),
                                hz = __shfl_sync(0xffffffffu, g0.z, src);
                    bool cand = false;
                    if (t < total) {
                        const float gx = fmaf(fi, p.step[0], hx), gy = fmaf(fi, p.step[1], hy), gz = fmaf(fi, p.step[2], hz);
                        const float z = fmaf(mz0, gx, fmaf(mz1, gy, mz2 * gz));
                        cand = coarse_test(z, fmaf(gx, gx, fmaf(gy, gy, gz * gz)), cc) && live_row(p, index);
                    }
                    append(cand, index);
                }
            }
        } else {
        auto test_block = [&](int i0, const float4 (&gk)[4]) {
            bool cands[4], any = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cands[k] = coarse_test(fmaf(mz0, gk[k].x, fmaf(mz1, gk[k].y, mz2 * gk[k].z)), gk[k].w, cc);
                any |= cands[k];
            }
            if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
            for (int k = 0; k < 4; ++k) append(cands[k], i0 + 32 * k + lane);
        };
        const int full_end = row_lo + ((row_hi - row_lo) & ~127);
        if (row_lo < full_end) {
            float4 nxt[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) nxt[k] = __ldg(p.g_f32 + row_lo + 32 * k + lane);
            for (int i0 = row_lo; i0 < full_end; i0 += 128) {
                float4 gk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) gk[k] = nxt[k];
                if (i0 + 128 < full_end) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) nxt[k] = __ldg(p.g_f32 + i0 + 128 + 32 * k + lane);
                }
                test_block(i0, gk);
            }
        }
        if (full_end < row_hi) {
            float4 gk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = full_end + 32 * k + lane;
                gk[k] = i < row_hi ? __ldg(p.g_f32 + i) : make_float4(0.f, 0.f, 0.f, INFINITY);
            }
            test_block(full_end, gk);
        }
        }
        if (lane == 0) s_ncand[warp] = full ? -1 : n_c;
        __syncthreads();
        bool overflow = false;
        for (int w = 0; w < SIM_WARPS; ++w) overflow |= s_ncand[w] < 0;
        if (overflow) {
            // ---- fall-back: warp 0 alone, the one-warp algorithm over the whole table
            if (warp == 0) {
                WarpState ws;
#pragma unroll
                for (int k = 0; k < 9; ++k) ws.m[k] = m[k];
                ws.n_out = 0;
                ws.max_I = -INFINITY;
                int n_list = 0;
                for (int i0 = 0; i0 < p.n_g; i0 += 32) {
                    const int i = i0 + lane;
                    bool cand = false;
                    if (i < p.n_g) {
                        const float4 g = __ldg(p.g_f32 + i);
                        cand = coarse_test(fmaf(mz0, g.x, fmaf(mz1, g.y, mz2 * g.z)), g.w, cc);
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, cand);
                    if (mask == 0u) continue;
                    if (cand) s_list[n_list + __popc(mask & ((1u << lane) - 1u))] = i;
                    n_list += __popc(mask);
                    __syncwarp();
                    if (n_list >= 32) {
                        refine<MODEL>(p, ws, rot, true, s_list[lane], lane, s_cos);
                        const int rest = n_list - 32;
                        const int carry = (lane < rest) ? s_list[32 + lane] : 0;
                        __syncwarp();
                        s_list[lane] = carry;
                        n_list = rest;
                        __syncwarp();
                    }
                }
                if (n_list > 0) refine<MODEL>(p, ws, rot, lane < n_list, lane < n_list ? s_list[lane] : 0, lane, s_cos);
                __syncwarp();
                local_max_count = max(local_max_count, ws.n_out);
                const int n_keep = threshold_row<MODEL>(p, rot, ws.n_out, ws.max_I, lane);
                if (lane == 0) p.count[rot] = n_keep;
            }
            continue;  // (uniform: the flags are in shared memory)
        }
        // ---- 2. float64 evaluation of this warp's candidates
        int n_out = 0;
        double max_I = -INFINITY;
        for (int j0 = 0; j0 < n_c; j0 += 32) {
            const int j = j0 + lane;
            const bool have = j < n_c;
            const int pos = have ? at(j) : 0;
            const Refined e = refine_eval<MODEL, true>(p, m, have, have ? s_cand[pos] : 0, lane, s_cos);
            if (have) s_I[pos] = e.keep ? e.I : __longlong_as_double(0x7ff8000000000000ll);
            n_out += __popc(__ballot_sync(0xffffffffu, e.keep));
            max_I = fmax(max_I, e.keep ? e.I : -INFINITY);
        }
        max_I = warp_max(max_I);
        if (lane == 0) {
            s_nout[warp] = n_out;
            s_wmax[warp] = max_I;
        }
        __syncthreads();
        int total_out = 0;
        double row_max = -INFINITY;
        for (int w = 0; w < SIM_WARPS; ++w) {
            total_out += s_nout[w];
            row_max = fmax(row_max, s_wmax[w]);
        }
        local_max_count = max(local_max_count, total_out);
        // ---- 3. survivors of the intensity cut per warp (NaN compares false; cut disabled: everything that is not NaN)
        const double cut = cut_off ? -INFINITY : row_max * p.min_intensity;
        int n_keep = 0;
        for (int j0 = 0; j0 < n_c; j0 += 32) {
            const int j = j0 + lane;
            const double I = j < n_c ? s_I[at(j)] : __longlong_as_double(0x7ff8000000000000ll);
            n_keep += __popc(__ballot_sync(0xffffffffu, cut_off ? I == I : I > cut));
        }
        if (lane == 0) s_nkeep[warp] = n_keep;
        __syncthreads();
        int offset = 0, total_keep = 0;
        for (int w = 0; w < SIM_WARPS; ++w) {
            if (w < warp) offset += s_nkeep[w];
            total_keep += s_nkeep[w];
        }
        // ---- 4. geometry again for the survivors, stored at their final positions
        for (int j0 = 0; j0 < n_c; j0 += 32) {
            const int j = j0 + lane;
            const int pos = j < n_c ? at(j) : 0;
            const double I = j < n_c ? s_I[pos] : __longlong_as_double(0x7ff8000000000000ll);
            const bool keep = cut_off ? I == I : I > cut;
            const int gi = keep ? s_cand[pos] : 0;
            const Refined e = refine_eval<MODEL, false>(p, m, keep, gi, lane, s_cos);
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            const int slot = offset + __popc(mask & ((1u << lane) - 1u));
            if (keep && slot < p.cap) {
                const size_t o = (size_t)rot * p.cap + slot;
                p.xyz[3 * o + 0] = e.x;
                p.xyz[3 * o + 1] = e.y;
                p.xyz[3 * o + 2] = e.z;
                p.intensity[o] = I;
                p.g_index[o] = gi;
                if (p.exc) p.exc[o] = e.s;
            }
            offset += __popc(mask);
        }
        if (threadIdx.x == 0) p.count[rot] = min(total_keep, p.cap);
    }
    local_max_count = warp_max(local_max_count);
    if (lane == 0 && local_max_count > 0) atomicMax(p.max_count, local_max_count);
}

// Rows whose |F|^2 can never pass the minimum-intensity cut (extinct reflections of centred / glide lattices,
// I0 <= rel_cut * I0[ref_row]) get |g|^2 = +inf: the coarse test then rejects them for free (see ds_pack_gtable
// in the header for why this leaves every result unchanged).
__global__ void pack_gtable_kernel(int n_g, const double *__restrict__ g, float4 *__restrict__ out,
                                   const double *__restrict__ I0, int ref_row, double rel_cut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_g) {
        const double x = g[3 * i], y = g[3 * i + 1], z = g[3 * i + 2];
        float g2 = (float)(x * x + y * y + z * z);
        if (I0 != nullptr && I0[i] <= rel_cut * I0[ref_row]) g2 = INFINITY;
        out[i] = make_float4((float)x, (float)y, (float)z, g2);
    }
}

}  // namespace ds

extern "C" int ds_pack_gtable(void *stream, int32_t n_g, const double *g_xyz, float *g_f32, const double *g_I0,
                              int32_t ref_row, double rel_cut) {
    using namespace ds;
    DS_REQUIRE(n_g >= 0, "ds_pack_gtable: negative size");
    if (n_g == 0) return 0;
    const bool mark = g_I0 != nullptr && ref_row >= 0 && rel_cut > 0.0;
    DS_REQUIRE(!mark || (ref_row < n_g && rel_cut < 1.0), "ds_pack_gtable: bad reference row / relative cut");
    pack_gtable_kernel<<<(n_g + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_g, g_xyz, reinterpret_cast<float4 *>(g_f32), mark ? g_I0 : nullptr, ref_row, rel_cut);
    return check_launch("ds_pack_gtable");
}

extern "C" int ds_simulate(void *stream, int32_t n_rot, const double *quat, int32_t n_g, const double *g_xyz,
                           const float *g_f32, const double *g_I0, double g_max, double inv_wavelength, double s_max,
                           double width, int32_t shape_model, double minima_number, double precession_rad,
                           double min_intensity, int32_t cap, int32_t *count, int32_t *g_index, double *xyz,
                           double *intensity, double *excitation_error, int32_t *max_count, int32_t n_lines,
                           const float *line_g0, const int32_t *line_start, const double *line_step_host) {
    using namespace ds;
    DS_REQUIRE(n_rot >= 0 && n_g >= 0 && cap > 0, "ds_simulate: bad sizes (n_rot=%d n_g=%d cap=%d)", n_rot, n_g,
               cap);
    DS_REQUIRE(shape_model >= 0 && shape_model <= DS_SHAPE_NONE_RETURN_S, "ds_simulate: unknown shape model %d",
               shape_model);
    DS_REQUIRE((reinterpret_cast<uintptr_t>(g_f32) & 15) == 0, "ds_simulate: g_f32 must be 16-byte aligned");
    DS_REQUIRE(inv_wavelength > 0, "ds_simulate: inv_wavelength must be positive");
    if (n_rot == 0) return 0;
    SimParams p;
    p.n_rot = n_rot;
    p.n_g = n_g;
    p.cap = cap;
    p.quat = quat;
    p.g_xyz = g_xyz;
    p.g_f32 = reinterpret_cast<const float4 *>(g_f32);
    p.g_I0 = g_I0;
    p.rs = inv_wavelength;
    p.s_max = s_max;
    p.width = width;
    p.minima = minima_number;
    p.prec = precession_rad;
    p.min_intensity = min_intensity;
    p.model = shape_model;
    p.count = count;
    p.g_index = g_index;
    p.xyz = xyz;
    p.intensity = intensity;
    p.exc = excitation_error;
    p.max_count = max_count;
    // float32 coarse cull margin: the error of s is dominated by the rounding of z' = R[2,:].g,
    // <~ 3e-7 |g|; 8e-6 max(1, |g|max) leaves > 20x head-room and admits < 0.1 % extra candidates.
    p.coarse_margin = 8e-6f * (float)(g_max > 1.0 ? g_max : 1.0);

    // Which kernel: large tables / few rotations take the CTA-per-rotation kernel (see below), everything else one warp per
    // rotation.  sim_cta = 0 never, 1 forces it.
    const int cta_opt = option(OPT_SIM_CTA);
    // default: tables too large to stay resident in the warp kernel's shared memory take it when they are very large
    // or the rotations few; resident tables (<= 6144 rows) only when a launch cannot fill the warp kernel's CTAs
    // (Fe3C at r = 2, 5 222 rows, 16 384 rotations: 160 us warp per rotation vs 271 us CTA per rotation)
    const bool cta_auto = n_g >= 4096 && (n_g > SIM_RESIDENT_MAX_G ? (n_g >= 32768 || n_rot <= 2048) : n_rot < 1024);
    const bool use_cta = n_g > 0 && (cta_opt > 0 || (cta_opt < 0 && cta_auto));
    // scan-line mode: the caller described the table as lattice lines (see include/diffsims_b200.h)
    size_t line_bytes = (size_t)n_lines * 16 + (size_t)((n_lines + 1 + 3) & ~3) * 4;
    bool lines = n_lines > 0 && line_g0 && line_start && line_step_host &&
                 (reinterpret_cast<uintptr_t>(line_g0) & 15) == 0 && (reinterpret_cast<uintptr_t>(line_start) & 15) == 0;
    if (lines && !use_cta) lines = line_bytes <= 96 * 1024;  // (the warp kernel keeps the line tables in shared memory)
    // DS_SIM_LINES = 0 / 1 overrides the measured rules (tests).
    if (lines) {
        const double rs = inv_wavelength, gm = g_max < rs ? g_max : rs * 0.999999;
        const double slab = 2.0 * s_max + rs - sqrt(rs * rs - gm * gm) + 2.0 * rs * sin(fabs(precession_rad)) * gm / rs;
        const double step = sqrt(line_step_host[0] * line_step_host[0] + line_step_host[1] * line_step_host[1] +
                                 line_step_host[2] * line_step_host[2]);
        const int force = option(OPT_SIM_LINES);
        if (force >= 0)
            lines = force != 0;
        else if (use_cta)
            // CTA kernel: a line at angle theta to the beam crosses the slab in slab / (step |cos theta|) rows, ~4.5 slab / step
            // on average over orientations; the interval expansion pays when that is a small part of a line
            lines = (double)n_g / n_lines >= 3.0 * (1.0 + 4.5 * slab / step);
        else
            // warp kernel (tools/bench_configs.py): solving per line only pays when the table is large and a line crosses
            // the slab in less than about half an index on average
            lines = n_g >= 2048 && slab < 0.5 * step;
    }
    p.n_lines = lines ? n_lines : 0;
    p.line_g0 = reinterpret_cast<const float4 *>(line_g0);
    p.line_start = line_start;
    p.z_lo = p.z_hi = 0.f;
    p.step[0] = p.step[1] = p.step[2] = 0.f;
    if (lines) {
        for (int k = 0; k < 3; ++k) p.step[k] = (float)line_step_host[k];
        // slab of z' = (R g).z that can hold a reflection: z_sphere(r) - t <= z' <= z_sphere(r) + t for r in
        // [0, g_max]; with precession the two tilted-sphere surfaces of simulation_generator.py:365-375
        const double rs = inv_wavelength, t = s_max + p.coarse_margin, gm = g_max < rs ? g_max : rs * 0.999999;
        double lo, hi;
        if (precession_rad == 0.0) {
            lo = -t;
            hi = rs - sqrt(rs * rs - gm * gm) + t;
        } else {
            const double Pz = rs * cos(precession_rad), Pt = rs * sin(precession_rad);
            const double rr = gm + Pt < rs ? gm + Pt : rs * 0.999999;
            lo = Pz - rs - t;                          // min over r of z_do(r), at r = P_t
            hi = Pz - sqrt(rs * rs - rr * rr) + t;     // max over r of z_up(r), at r = g_max
        }
        p.z_lo = (float)(lo - p.coarse_margin);
        p.z_hi = (float)(hi + p.coarse_margin);
    }

    const int n_tiles = lines ? 1 : ((n_g <= SIM_RESIDENT_MAX_G) ? 1 : (n_g + SIM_TILE_G - 1) / SIM_TILE_G);
    const int tile_g = lines ? (int)((line_bytes + 15) / 16) : ((n_tiles == 1) ? (n_g > 0 ? n_g : 1) : SIM_TILE_G);
    // precession with a model other than the closed-form Lorentzian: numerical average over the circle
    p.n_quad = 0;
    if (precession_rad != 0.0 && shape_model != DS_SHAPE_LORENTZIAN_PRECESSION &&
        shape_model != DS_SHAPE_NONE_RETURN_S && shape_model != DS_SHAPE_BINARY)
        p.n_quad = (shape_model == DS_SHAPE_LORENTZIAN || shape_model == DS_SHAPE_ATANC) ? 2048 : 8192;
    const int tile_alloc = lines ? tile_g : ((tile_g + 127) & ~127);
    const size_t smem = (size_t)tile_alloc * 16 * (n_tiles > 1 ? 2 : 1) + (size_t)p.n_quad * 8;
    // Few rotations over a large table (one warp per rotation cannot fill 148 SMs): smaller CTAs spread the rotations
    // over more SMs; the table is then streamed by more CTAs, which L2 absorbs.  sim_split = 1 / 2 / 4 / 8 forces the
    // warps per CTA.
    // Measured on the 113 082-row table (tools/bench_k12_large.py): 512 rotations 351 / 308 / 488 us with 8 / 2 / 1 warps
    // per CTA, 2 048 rotations 428 / 993 / 1535 us.
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Large tables: one CTA per rotation, the table split across its warps (simulate_cta_kernel).  Measured on the
    // 113 082-row table (tools/bench_k12_large.py, profiles/r02_k12_large.txt), CTA per rotation vs warp per rotation:
    // 128 rotations 123 vs 322 us, 512: 138 vs 329 us, 2 048: 337 vs 403 us, 16 384: 2 227 vs 2 434 us -- the CTAs of an SM
    // walk the table in step and share its lines in L1, and nothing is staged through shared memory.  (24 406 rows: equal at
    // 16 384 rotations, 81 vs 131 us at 512.)  sim_cta = 0 never, 1 forces it; sim_stash = candidate capacity of a rotation (small values exercise the fall-back in the tests).
    {
        if (use_cta) {
            int stash = option(OPT_SIM_STASH);
            if (stash < SIM_CHUNK) stash = 4096;
            int n_chunks = (stash + SIM_CHUNK - 1) / SIM_CHUNK;
            if (n_chunks > SIM_MAX_CHUNKS) n_chunks = SIM_MAX_CHUNKS;
            const size_t smem_c = (size_t)n_chunks * SIM_CHUNK * 12 + (((size_t)SIM_WARPS * n_chunks + 15) & ~(size_t)15) +
                                  (size_t)p.n_quad * 8;
            auto launch_cta = [&](auto kern) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
                int blocks_per_sm = 1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, SIM_THREADS, smem_c);
                if (blocks_per_sm < 1) blocks_per_sm = 1;
                const int grid = n_rot < num_sms() * blocks_per_sm ? n_rot : num_sms() * blocks_per_sm;
                kern<<<grid, SIM_THREADS, smem_c, st>>>(p, n_chunks);
            };
#define DS_SIM_CTA(M)                                     \
    do {                                                  \
        if (lines)                                        \
            launch_cta(simulate_cta_kernel<M, true>);     \
        else                                              \
            launch_cta(simulate_cta_kernel<M, false>);    \
    } while (0)
            if (precession_rad != 0.0) {
                DS_SIM_CTA(-1);
            } else {
                switch (shape_model) {
                    case DS_SHAPE_BINARY: DS_SIM_CTA(DS_SHAPE_BINARY); break;
                    case DS_SHAPE_LINEAR: DS_SIM_CTA(DS_SHAPE_LINEAR); break;
                    case DS_SHAPE_SINC: DS_SIM_CTA(DS_SHAPE_SINC); break;
                    case DS_SHAPE_SIN2C: DS_SIM_CTA(DS_SHAPE_SIN2C); break;
                    case DS_SHAPE_ATANC: DS_SIM_CTA(DS_SHAPE_ATANC); break;
                    case DS_SHAPE_LORENTZIAN: DS_SIM_CTA(DS_SHAPE_LORENTZIAN); break;
                    case DS_SHAPE_NONE_RETURN_S: DS_SIM_CTA(DS_SHAPE_NONE_RETURN_S); break;
                    default: DS_SIM_CTA(-1); break;  // lorentzian_precession with zero angle
                }
            }
#undef DS_SIM_CTA
            return check_launch("ds_simulate (CTA per rotation)");
        }
    }
    int wpb = SIM_WARPS;
    if (n_g >= 2048 && n_rot < 1024) wpb = 2;
    {
        const int o = option(OPT_SIM_SPLIT);
        if (o == 1 || o == 2 || o == 4 || o == 8) wpb = o;
    }
    const int n_batches = (n_rot + wpb - 1) / wpb;
    // no precession: one lean kernel per shape factor model; anything with precession: the general kernel
    auto launch = [&](auto kern, int slot) {
        (void)slot;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SIM_TILE_G * 16 + 8192 * 8);
        int blocks_per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, wpb * 32, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        const int grid = n_batches < num_sms() * blocks_per_sm ? n_batches : num_sms() * blocks_per_sm;
        kern<<<grid, wpb * 32, smem, st>>>(p, n_tiles, tile_g, tile_alloc);
    };
#define DS_SIM(M, SLOT)                                   \
    do {                                                  \
        if (lines)                                        \
            launch(simulate_kernel<M, true>, 9 + SLOT);   \
        else                                              \
            launch(simulate_kernel<M, false>, SLOT);      \
    } while (0)
    if (precession_rad != 0.0) {
        DS_SIM(-1, 8);
    } else {
        switch (shape_model) {
            case DS_SHAPE_BINARY: DS_SIM(DS_SHAPE_BINARY, 0); break;
            case DS_SHAPE_LINEAR: DS_SIM(DS_SHAPE_LINEAR, 1); break;
            case DS_SHAPE_SINC: DS_SIM(DS_SHAPE_SINC, 2); break;
            case DS_SHAPE_SIN2C: DS_SIM(DS_SHAPE_SIN2C, 3); break;
            case DS_SHAPE_ATANC: DS_SIM(DS_SHAPE_ATANC, 4); break;
            case DS_SHAPE_LORENTZIAN: DS_SIM(DS_SHAPE_LORENTZIAN, 5); break;
            case DS_SHAPE_NONE_RETURN_S: DS_SIM(DS_SHAPE_NONE_RETURN_S, 7); break;
            default: DS_SIM(-1, 8); break;  // lorentzian_precession with zero angle
        }
    }
#undef DS_SIM
    return check_launch("ds_simulate");
}
