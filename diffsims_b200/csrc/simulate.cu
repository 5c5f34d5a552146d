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

struct WarpState {
    // float64 active rotation matrix, row-major
    double m[9];
    int n_out;     // reflections that passed the excitation-error test so far
    double max_I;  // running max of their intensities
};

// Refine up to 32 candidates (one per lane) in float64 and append the survivors to the output row.
// MODEL >= 0 fixes the shape factor at compile time (no precession); MODEL < 0 is the general kernel
// (runtime model, precession cut, closed-form or numerically averaged precession shape factor).
template <int MODEL>
__device__ __forceinline__ void refine(const SimParams &p, WarpState &w, int rot, bool have, int gi, int lane,
                                       const double *__restrict__ s_cos) {
    constexpr bool GENERAL = MODEL < 0;
    const int model = GENERAL ? p.model : MODEL;
    bool keep = false;
    double x = 0, y = 0, z = 0, s = 0, I = 0, r_spot = 0;
    if (have) {
        const double gx = __ldg(p.g_xyz + 3 * (size_t)gi), gy = __ldg(p.g_xyz + 3 * (size_t)gi + 1),
                     gz = __ldg(p.g_xyz + 3 * (size_t)gi + 2);
        x = w.m[0] * gx + w.m[1] * gy + w.m[2] * gz;
        y = w.m[3] * gx + w.m[4] * gy + w.m[5] * gz;
        z = w.m[6] * gx + w.m[7] * gy + w.m[8] * gz;
        // simulation_generator.py:355-360, evaluated as the reference writes it
        r_spot = sqrt(x * x + y * y);
        const double z_sphere = -sqrt(p.rs * p.rs - r_spot * r_spot) + p.rs;
        s = z_sphere - z;
        if (!GENERAL || p.prec == 0.0) {
            keep = fabs(s) < p.s_max;  // :364 strict
        } else {                       // :365-375
            const double P_z = p.rs * cos(p.prec), P_t = p.rs * sin(p.prec);
            const double up = P_z - sqrt(p.rs * p.rs - (r_spot + P_t) * (r_spot + P_t));
            const double dn = P_z - sqrt(p.rs * p.rs - (r_spot - P_t) * (r_spot - P_t));
            keep = (z - p.s_max <= up) && (z + p.s_max >= dn);
        }
    }
    double sf = 1.0;
    if (GENERAL && p.n_quad > 0) {
        // _shape_factor_precession (shape_factor_models.py:222-269): (1 / 2 pi) int_0^2pi f(s + r phi cos t) dt.
        // The integrand is even and periodic in t, so the midpoint rule on [0, pi] (Gauss-Chebyshev in
        // u = cos t) converges geometrically for the smooth models and as 1/n^2 for the kinked ones; the
        // whole warp integrates one candidate at a time over a cosine table in shared memory.
        for (int c = 0; c < 32; ++c) {
            if (!__shfl_sync(0xffffffffu, (int)keep, c)) continue;
            const double sc = __shfl_sync(0xffffffffu, s, c);
            const double amp = __shfl_sync(0xffffffffu, r_spot, c) * p.prec;
            double acc = 0.0;
            for (int j = lane; j < p.n_quad; j += 32)
                acc += shape_factor(model, sc + amp * s_cos[j], p.width, p.minima, 0.0, 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == c) sf = acc / (double)p.n_quad;
        }
    } else if (keep) {
        sf = shape_factor(model, s, p.width, p.minima, r_spot, GENERAL ? p.prec : 0.0);
    }
    if (keep) I = sf * __ldg(p.g_I0 + gi);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int slot = w.n_out + __popc(mask & ((1u << lane) - 1u));
    if (keep && slot < p.cap) {
        const size_t o = (size_t)rot * p.cap + slot;
        p.xyz[3 * o + 0] = x;
        p.xyz[3 * o + 1] = y;
        p.xyz[3 * o + 2] = z;
        p.intensity[o] = I;
        p.g_index[o] = gi;
        if (p.exc) p.exc[o] = s;
    }
    w.n_out += __popc(mask);
    w.max_I = fmax(w.max_I, warp_max(keep ? I : -INFINITY));
}

template <int MODEL>
__global__ void __launch_bounds__(SIM_THREADS, MODEL < 0 ? 2 : 3) simulate_kernel(const SimParams p, const int n_tiles,
                                                                  const int tile_g) {
    constexpr bool GENERAL = MODEL < 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s_tile[2] = {reinterpret_cast<float4 *>(smem_raw),
                         reinterpret_cast<float4 *>(smem_raw) + (n_tiles > 1 ? tile_g : 0)};
    double *s_cos = reinterpret_cast<double *>(smem_raw + (size_t)tile_g * 16 * (n_tiles > 1 ? 2 : 1));
    if (GENERAL)
        for (int j = threadIdx.x; j < p.n_quad; j += SIM_THREADS) s_cos[j] = cospi(((double)j + 0.5) / (double)p.n_quad);
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

    if (n_tiles == 1) {  // resident table: one bulk copy for the CTA's lifetime
        issue(0, 0);
        mbar_wait(&s_bar[0], 0);
    }

    const float two_rs = 2.0f * (float)p.rs;
    const float thr = (float)p.s_max + p.coarse_margin;
    const bool prec_on = GENERAL && p.prec != 0.0;
    const float P_z = GENERAL ? (float)(p.rs * cos(p.prec)) : 0.f, P_t = GENERAL ? (float)(p.rs * sin(p.prec)) : 0.f;
    const uint32_t tile_base_s = smem_u32(smem_raw);
    int local_max_count = 0;

    const int n_batches = (p.n_rot + SIM_WARPS - 1) / SIM_WARPS;
    for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        const int rot = batch * SIM_WARPS + warp;
        const bool active = rot < p.n_rot;
        WarpState w;
        w.n_out = 0;
        w.max_I = -INFINITY;
        float mz0 = 0, mz1 = 0, mz2 = 0;
        if (active) {
            const double a = p.quat[4 * (size_t)rot], b = p.quat[4 * (size_t)rot + 1],
                         c = p.quat[4 * (size_t)rot + 2], d = p.quat[4 * (size_t)rot + 3];
            w.m[0] = a * a + b * b - c * c - d * d;
            w.m[1] = 2 * (b * c - a * d);
            w.m[2] = 2 * (b * d + a * c);
            w.m[3] = 2 * (b * c + a * d);
            w.m[4] = a * a - b * b + c * c - d * d;
            w.m[5] = 2 * (c * d - a * b);
            w.m[6] = 2 * (b * d - a * c);
            w.m[7] = 2 * (c * d + a * b);
            w.m[8] = a * a - b * b - c * c + d * d;
            mz0 = (float)w.m[6];
            mz1 = (float)w.m[7];
            mz2 = (float)w.m[8];
        }
        int n_list = 0;
        int *list = s_list[warp];
        const uint32_t list_s = smem_u32(list);

        if (n_tiles > 1) issue(0, 0);
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = (n_tiles > 1) ? (t & 1) : 0;
            if (n_tiles > 1) {
                if (t + 1 < n_tiles) issue(t + 1, (t + 1) & 1);
                mbar_wait(&s_bar[buf], phase[buf]);
                phase[buf] ^= 1;
            }
            const int n = min(tile_g, p.n_g - t * tile_g);
            const uint32_t tile_s = tile_base_s + (buf ? (uint32_t)tile_g * 16u : 0u);
            if (active) {
                for (int i0 = 0; i0 < n; i0 += 128) {
                    // four independent g per lane: loads and tests overlap, ballots are consumed in table order
                    unsigned masks[4];
                    bool cands[4];
                    float4 gk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        gk[k] = lds_f4(tile_s + 16u * (uint32_t)min(i0 + 32 * k + lane, n - 1));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = i0 + 32 * k + lane;
                        const float4 g = gk[k];
                        const float z = fmaf(mz0, g.x, fmaf(mz1, g.y, mz2 * g.z));
                        const float r2 = fmaf(-z, z, g.w);
                        const float u = z + thr, v = z - thr;
                        bool c;
                        if (!prec_on) {
                            // |s| < t  <=>  f(z + t) < 0 < f(z - t),  f(u) = r^2 + u (u - 2 r_s)
                            c = (fmaf(u, u - two_rs, r2) < 0.0f) && (fmaf(v, v - two_rs, r2) > 0.0f);
                        } else {
                            // z - t <= z_up(r) and z + t >= z_do(r) (simulation_generator.py:365-375), same algebra
                            // with the tilted sphere centre (P_t, P_z)
                            const float r = sqrtf(fmaxf(r2, 0.0f)), two_rpt = 2.0f * r * P_t;
                            c = (r2 + two_rpt + v * (v - 2.0f * P_z) >= 0.0f) && (r2 - two_rpt + u * (u - 2.0f * P_z) <= 0.0f);
                        }
                        cands[k] = c && (i < n);
                        masks[k] = __ballot_sync(0xffffffffu, cands[k]);
                    }
                    if ((masks[0] | masks[1] | masks[2] | masks[3]) == 0u) continue;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned mask = masks[k];
                        if (mask == 0u) continue;
                        if (cands[k])
                            sts_u32(list_s + 4u * (uint32_t)(n_list + __popc(mask & ((1u << lane) - 1u))),
                                    (uint32_t)(t * tile_g + i0 + 32 * k + lane));
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
                    }
                }
            }
            if (n_tiles > 1) __syncthreads();  // everyone is done with `buf` before it is refilled
        }
        if (active) {
            if (n_list > 0) refine<MODEL>(p, w, rot, lane < n_list, lane < n_list ? list[lane] : 0, lane, s_cos);
            __syncwarp();
            local_max_count = max(local_max_count, w.n_out);
            // ---- threshold: keep I > max(I) * min_intensity (simulation_generator.py:237), in place
            const int n_stored = min(w.n_out, p.cap);
            int n_keep = 0;
            if ((GENERAL ? p.model : MODEL) == DS_SHAPE_NONE_RETURN_S || p.min_intensity < 0.0) {  // threshold disabled
                n_keep = n_stored;
            } else {
                const double cut = w.max_I * p.min_intensity;
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
            }
            if (lane == 0) p.count[rot] = n_keep;
        }
    }
    local_max_count = warp_max(local_max_count);
    if (lane == 0 && local_max_count > 0) atomicMax(p.max_count, local_max_count);
}

__global__ void pack_gtable_kernel(int n_g, const double *__restrict__ g, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_g) {
        const double x = g[3 * i], y = g[3 * i + 1], z = g[3 * i + 2];
        out[i] = make_float4((float)x, (float)y, (float)z, (float)(x * x + y * y + z * z));
    }
}

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

}  // namespace ds

extern "C" int ds_pack_gtable(void *stream, int32_t n_g, const double *g_xyz, float *g_f32) {
    using namespace ds;
    DS_REQUIRE(n_g >= 0, "ds_pack_gtable: negative size");
    if (n_g == 0) return 0;
    pack_gtable_kernel<<<(n_g + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_g, g_xyz, reinterpret_cast<float4 *>(g_f32));
    return check_launch("ds_pack_gtable");
}

extern "C" int ds_simulate(void *stream, int32_t n_rot, const double *quat, int32_t n_g, const double *g_xyz,
                           const float *g_f32, const double *g_I0, double g_max, double inv_wavelength, double s_max,
                           double width, int32_t shape_model, double minima_number, double precession_rad,
                           double min_intensity, int32_t cap, int32_t *count, int32_t *g_index, double *xyz,
                           double *intensity, double *excitation_error, int32_t *max_count) {
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

    const int n_tiles = (n_g <= SIM_RESIDENT_MAX_G) ? 1 : (n_g + SIM_TILE_G - 1) / SIM_TILE_G;
    const int tile_g = (n_tiles == 1) ? (n_g > 0 ? n_g : 1) : SIM_TILE_G;
    // precession with a model other than the closed-form Lorentzian: numerical average over the circle
    p.n_quad = 0;
    if (precession_rad != 0.0 && shape_model != DS_SHAPE_LORENTZIAN_PRECESSION &&
        shape_model != DS_SHAPE_NONE_RETURN_S && shape_model != DS_SHAPE_BINARY)
        p.n_quad = (shape_model == DS_SHAPE_LORENTZIAN || shape_model == DS_SHAPE_ATANC) ? 2048 : 8192;
    const size_t smem = (size_t)tile_g * 16 * (n_tiles > 1 ? 2 : 1) + (size_t)p.n_quad * 8;
    const int n_batches = (n_rot + SIM_WARPS - 1) / SIM_WARPS;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // no precession: one lean kernel per shape factor model; anything with precession: the general kernel
    auto launch = [&](auto kern, int slot) {
        static bool attr_set[9] = {false};
        if (!attr_set[slot]) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SIM_TILE_G * 16 + 8192 * 8);
            attr_set[slot] = true;
        }
        int blocks_per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, SIM_THREADS, smem);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        const int grid = n_batches < num_sms() * blocks_per_sm ? n_batches : num_sms() * blocks_per_sm;
        kern<<<grid, SIM_THREADS, smem, st>>>(p, n_tiles, tile_g);
    };
    if (precession_rad != 0.0) {
        launch(simulate_kernel<-1>, 8);
    } else {
        switch (shape_model) {
            case DS_SHAPE_BINARY: launch(simulate_kernel<DS_SHAPE_BINARY>, 0); break;
            case DS_SHAPE_LINEAR: launch(simulate_kernel<DS_SHAPE_LINEAR>, 1); break;
            case DS_SHAPE_SINC: launch(simulate_kernel<DS_SHAPE_SINC>, 2); break;
            case DS_SHAPE_SIN2C: launch(simulate_kernel<DS_SHAPE_SIN2C>, 3); break;
            case DS_SHAPE_ATANC: launch(simulate_kernel<DS_SHAPE_ATANC>, 4); break;
            case DS_SHAPE_LORENTZIAN: launch(simulate_kernel<DS_SHAPE_LORENTZIAN>, 5); break;
            case DS_SHAPE_NONE_RETURN_S: launch(simulate_kernel<DS_SHAPE_NONE_RETURN_S>, 7); break;
            default: launch(simulate_kernel<-1>, 8); break;  // lorentzian_precession with zero angle
        }
    }
    return check_launch("ds_simulate");
}
