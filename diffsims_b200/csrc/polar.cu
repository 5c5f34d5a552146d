// polar_flatten_simulations on the packed result.
//
// Replaces the per-template Python loop of Simulation2D.polar_flatten_simulations
// (diffsims/simulations/simulation2d.py:313-355): r = |g_xy|, theta = atan2(y, x)
// (DiffractingVector.to_flat_polar, crystallography/_diffracting_vector.py:186-194), optional snapping to
// radial / azimuthal axes with get_closest (simulation2d.py:767-781), removal of out-of-range spots, zero
// padding to the longest template.  One warp per template, ballot compaction, float64 throughout.
#include "common.cuh"

namespace ds {

// numpy.searchsorted(side="left") followed by the "previous entry is closer" correction of get_closest
__device__ __forceinline__ int get_closest(const double *__restrict__ a, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    int idx = lo;
    const bool prev = (idx == n) || (fabs(v - a[max(idx - 1, 0)]) < fabs(v - a[min(idx, n - 1)]));
    if (prev) idx -= 1;
    return idx;
}

__global__ void polar_flatten_kernel(int n_tmpl, int cap, const int *__restrict__ count,
                                     const double *__restrict__ xyz, const double *__restrict__ intensity,
                                     int max_spots, int n_rad, const double *__restrict__ rad, int n_az,
                                     const double *__restrict__ az, double *__restrict__ r_out,
                                     double *__restrict__ t_out, double *__restrict__ i_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_tmpl) return;
    const int n = min(count[warp], cap);
    const size_t row = (size_t)warp * cap, orow = (size_t)warp * max_spots;
    const bool snap = rad != nullptr && az != nullptr;
    int n_out = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        bool keep = false;
        double r = 0, t = 0, I = 0;
        if (j < n) {
            const double x = xyz[3 * (row + j)], y = xyz[3 * (row + j) + 1];
            r = sqrt(x * x + y * y);
            t = atan2(y, x);
            I = intensity[row + j];
            keep = true;
            if (snap) {
                const int ri = get_closest(rad, n_rad, r), ti = get_closest(az, n_az, t);
                keep = (ri < n_rad - 1) && (ti < n_az - 1);
                r = (double)ri;
                t = (double)ti;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int d = n_out + __popc(mask & ((1u << lane) - 1u));
        if (keep && d < max_spots) {
            r_out[orow + d] = r;
            t_out[orow + d] = t;
            i_out[orow + d] = I;
        }
        n_out += __popc(mask);
    }
    for (int d = min(n_out, max_spots) + lane; d < max_spots; d += 32) {
        r_out[orow + d] = 0.0;
        t_out[orow + d] = 0.0;
        i_out[orow + d] = 0.0;
    }
}

// Library pixel coordinates, diffsims/generators/library_generator.py:129-132 with
// DiffractionSimulation.calibrated_coordinates (diffsims/sims/diffraction_simulation.py:143-149):
// rint((xy + offset) / calibration + half_shape) as int32; rint is round-half-to-even like numpy's.
__global__ void pixel_coords_kernel(int n_tmpl, int cap, const int *__restrict__ count,
                                    const double *__restrict__ xyz, double cal_x, double cal_y, double off_x,
                                    double off_y, double half_x, double half_y, int *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_tmpl * cap) return;
    const int t = (int)(i / cap), j = (int)(i % cap);
    int px = 0, py = 0;
    if (j < min(count[t], cap)) {
        px = (int)rint((xyz[3 * i] + off_x) / cal_x + half_x);
        py = (int)rint((xyz[3 * i + 1] + off_y) / cal_y + half_y);
    }
    out[2 * i] = px;
    out[2 * i + 1] = py;
}

// Padded rows -> CSR: template t's count[t] reflections go to offsets[t] .. offsets[t + 1] - 1 (the "packed spot
// lists" a sharded library gathers and a host consumer receives: ~40 bytes per reflection instead of `cap` slots).
__global__ void pack_csr_kernel(int n_tmpl, int cap, const int *__restrict__ count, const long long *__restrict__ offsets,
                                const int *__restrict__ g_index, const double *__restrict__ xyz,
                                const double *__restrict__ intensity, int *__restrict__ g_out, double *__restrict__ xyz_out,
                                double *__restrict__ i_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_tmpl) return;
    const int n = min(count[warp], cap);
    const size_t row = (size_t)warp * cap;
    const long long o = offsets[warp];
    if (g_out != nullptr)
        for (int j = lane; j < n; j += 32) g_out[o + j] = g_index[row + j];
    if (i_out != nullptr)
        for (int j = lane; j < n; j += 32) i_out[o + j] = intensity[row + j];
    if (xyz_out != nullptr)
        for (int j = lane; j < 3 * n; j += 32) xyz_out[3 * o + j] = xyz[3 * row + j];
}

}  // namespace ds

extern "C" int ds_pack_csr(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count, const int64_t *offsets,
                           const int32_t *g_index, const double *xyz, const double *intensity, int32_t *g_index_out,
                           double *xyz_out, double *intensity_out) {
    using namespace ds;
    DS_REQUIRE(n_tmpl >= 0 && cap > 0, "ds_pack_csr: bad sizes");
    DS_REQUIRE(count != nullptr && offsets != nullptr, "ds_pack_csr: count and offsets are required");
    DS_REQUIRE((g_index_out == nullptr || g_index != nullptr) && (xyz_out == nullptr || xyz != nullptr) &&
                   (intensity_out == nullptr || intensity != nullptr),
               "ds_pack_csr: an output was requested without its input");
    if (n_tmpl == 0) return 0;
    const int threads = 256, warps = threads / 32;
    pack_csr_kernel<<<(n_tmpl + warps - 1) / warps, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        n_tmpl, cap, count, reinterpret_cast<const long long *>(offsets), g_index, xyz, intensity, g_index_out, xyz_out,
        intensity_out);
    return check_launch("ds_pack_csr");
}

extern "C" int ds_library_pixel_coords(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count,
                                       const double *xyz, double calibration_x, double calibration_y,
                                       double offset_x, double offset_y, double half_shape_x, double half_shape_y,
                                       int32_t *pixel_coords) {
    using namespace ds;
    DS_REQUIRE(n_tmpl >= 0 && cap > 0, "ds_library_pixel_coords: bad sizes");
    DS_REQUIRE(calibration_x != 0.0 && calibration_y != 0.0, "ds_library_pixel_coords: calibration cannot be zero");
    if (n_tmpl == 0) return 0;
    const long long total = (long long)n_tmpl * cap;
    pixel_coords_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_tmpl, cap, count, xyz, calibration_x, calibration_y, offset_x, offset_y, half_shape_x, half_shape_y,
        pixel_coords);
    return check_launch("ds_library_pixel_coords");
}

extern "C" int ds_polar_flatten(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count, const double *xyz,
                                const double *intensity, int32_t max_spots, int32_t n_radial,
                                const double *radial_axes, int32_t n_azimuthal, const double *azimuthal_axes,
                                double *r_out, double *theta_out, double *intensity_out) {
    using namespace ds;
    DS_REQUIRE(n_tmpl >= 0 && cap > 0 && max_spots >= 0, "ds_polar_flatten: bad sizes");
    DS_REQUIRE((radial_axes == nullptr) == (azimuthal_axes == nullptr),
               "ds_polar_flatten: radial and azimuthal axes must be given together");
    if (n_tmpl == 0 || max_spots == 0) return 0;
    const int threads = 256, warps = threads / 32;
    polar_flatten_kernel<<<(n_tmpl + warps - 1) / warps, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        n_tmpl, cap, count, xyz, intensity, max_spots, n_radial, radial_axes, n_azimuthal, azimuthal_axes, r_out,
        theta_out, intensity_out);
    return check_launch("ds_polar_flatten");
}

// ---------------------------------------------------------------------------------------------------
// Optional 16-bit export of normalised templates: u = rint(clamp(v, 0, 1) * 65535).  The quantisation step is
// 1.5e-5 of the peak (the parity budget of the templates is 1e-4); host consumers that accept it halve the
// device->host bytes, which is what bounds the end-to-end image path.  Streaming kernel: 32 bytes in, 16 out per thread.
// ---------------------------------------------------------------------------------------------------
namespace ds {
__global__ void __launch_bounds__(256) quantize_u16_kernel(long long n8, long long n, const float *__restrict__ in,
                                                            unsigned short *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(in) + 2 * i), b = __ldcs(reinterpret_cast<const float4 *>(in) + 2 * i + 1);
        auto q = [](float v) { return (unsigned)__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * 65535.f); };
        uint4 o;
        o.x = q(a.x) | (q(a.y) << 16);
        o.y = q(a.z) | (q(a.w) << 16);
        o.z = q(b.x) | (q(b.y) << 16);
        o.w = q(b.z) | (q(b.w) << 16);
        __stcs(reinterpret_cast<uint4 *>(out) + i, o);
    }
    // tail (n not a multiple of 8)
    if (blockIdx.x == 0)
        for (long long i = 8 * n8 + threadIdx.x; i < n; i += blockDim.x)
            out[i] = (unsigned short)__float2int_rn(fminf(fmaxf(in[i], 0.f), 1.f) * 65535.f);
}
}  // namespace ds

extern "C" int ds_quantize_u16(void *stream, int64_t n, const float *in, uint16_t *out) {
    using namespace ds;
    DS_REQUIRE(n >= 0, "ds_quantize_u16: negative size");
    DS_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "ds_quantize_u16: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    const long long n8 = n / 8;
    long long blocks = (n8 + 255) / 256;
    const long long cap = 16ll * num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    quantize_u16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(n8, n, in, out);
    return check_launch("ds_quantize_u16");
}
