// Rotation-list producers over SO(3) on the device: an equal-volume cubochoric grid of
// rotations cropped to the fundamental zone of a proper point group, or to a neighbourhood of a centre rotation.
//
// Replaces get_fundamental_zone_grid / get_local_grid (diffsims/generators/rotation_list_generators.py:85-134), which
// call orix.sampling.get_sample_fundamental / get_sample_local.  orix is a third-party dependency whose source is not
// under /root/reference, so PARITY WITH ITS POINT LISTS IS UNPINNED; the kernel restates the published algorithm orix
// implements (Rosca, Morawiec, De Graef, "A new method of constructing a grid in the space of 3D rotations and its
// applications to texture analysis", Modelling Simul. Mater. Sci. Eng. 22 (2014) 075013; the cubochoric sampling of
// Singh and De Graef, ibid. 24 (2016) 085013):
//   cube of edge pi^(2/3), (2 N)^3 cell-centred points  (x = (i - 1/2) delta, i = -N + 1 .. N, delta = pi^(2/3) / (2 N))
//   -> equal-volume map onto the homochoric ball (radius (3 pi / 4)^(1/3)) -> axis-angle (|h|^3 = 3/4 (w - sin w),
//   solved by Newton iteration instead of the usual polynomial fit) -> unit quaternion with non-negative scalar part.
// Crops:
//   mode 0  none (the whole of SO(3));
//   mode 1  fundamental zone of a proper point group given as its n_sym quaternions: q is kept iff no symmetric
//           equivalent s q has a smaller rotation angle, i.e. q.w >= |(s q).w| - tol for all s (the Rodrigues-space
//           region orix's OrientationRegion.from_symmetry describes);
//   mode 2  rotation angle <= max_angle (a ball around the identity), the local grid.
// Kept rotations are optionally composed with a centre (q_out = centre * q) and written in grid order (two-pass
// ordered compaction as in ds_beam_grid) as Bunge Euler angles in degrees and / or as the ACTIVE quaternions
// ds_simulate consumes (the reference rotates g by ~rotation, crystallography/_diffracting_vector.py:160).
#include "common.cuh"

namespace ds {

constexpr int SO3_THREADS = 256;
constexpr int SO3_MAX_SYM = 24;

struct So3Params {
    int n_steps;  // N: semi-edge steps; 2 N points per cube edge
    long long n_points;
    int mode;
    int n_sym;
    double sym[SO3_MAX_SYM][4];
    double tol;
    double max_angle;
    double centre[4];
    int has_centre;
};

// cube (edge pi^(2/3)) -> homochoric ball, the inverse Lambert construction of the paper (section 3)
__device__ __forceinline__ void cubochoric_to_homochoric(double x, double y, double z, double &hx, double &hy, double &hz) {
    const double PI = 3.141592653589793;
    const double sc = 0.897772786961286;    // pi^(5/6) / 6^(1/6) / pi^(2/3)
    const double prek = 1.6434564029725040; // R1 2^(1/4) / beta
    const double pref = 1.3819765978853418; // sqrt(6 / pi)
    const double r2 = 1.4142135623730951, r24 = 4.898979485566356, spi = 1.7724538509055159;
    const double ax = fabs(x), ay = fabs(y), az = fabs(z);
    if (fmax(ax, fmax(ay, az)) == 0.0) {
        hx = hy = hz = 0.0;
        return;
    }
    // pyramid: permute so that the third coordinate is the dominant one
    int pyr;  // 0: +-z, 1: +-x, 2: +-y
    double a, b, c;
    if (ax <= az && ay <= az) {
        pyr = 0, a = x, b = y, c = z;
    } else if (ay <= ax && az <= ax) {
        pyr = 1, a = y, b = z, c = x;
    } else {
        pyr = 2, a = z, b = x, c = y;
    }
    a *= sc, b *= sc, c *= sc;
    double la, lb, lc;
    if (fmax(fabs(a), fabs(b)) == 0.0) {
        la = lb = 0.0;
        lc = pref * c;
    } else {
        double t1, t2;
        if (fabs(b) <= fabs(a)) {
            const double q = (PI / 12.0) * b / a;
            double s, co;
            sincos(q, &s, &co);
            const double f = prek * a / sqrt(r2 - co);
            t1 = (r2 * co - 1.0) * f;
            t2 = r2 * s * f;
        } else {
            const double q = (PI / 12.0) * a / b;
            double s, co;
            sincos(q, &s, &co);
            const double f = prek * b / sqrt(r2 - co);
            t1 = r2 * s * f;
            t2 = (r2 * co - 1.0) * f;
        }
        const double cc = t1 * t1 + t2 * t2;
        const double s = PI * cc / (24.0 * c * c);
        const double d = spi * cc / r24 / c;
        const double q = sqrt(1.0 - s);
        la = t1 * q, lb = t2 * q, lc = pref * c - d;
    }
    if (pyr == 0) {
        hx = la, hy = lb, hz = lc;
    } else if (pyr == 1) {
        hx = lc, hy = la, hz = lb;
    } else {
        hx = lb, hy = lc, hz = la;
    }
}

// homochoric vector -> unit quaternion (a >= 0): |h|^3 = 3/4 (w - sin w)
__device__ __forceinline__ void homochoric_to_quat(double hx, double hy, double hz, double (&q)[4]) {
    const double h2 = hx * hx + hy * hy + hz * hz;
    if (h2 == 0.0) {
        q[0] = 1.0, q[1] = q[2] = q[3] = 0.0;
        return;
    }
    const double h = sqrt(h2), target = (4.0 / 3.0) * h2 * h;  // w - sin w
    double w = cbrt(6.0 * target);                            // w - sin w ~ w^3 / 6
    w = fmin(w, 3.141592653589793);
    for (int it = 0; it < 12; ++it) {
        const double f = w - sin(w) - target, fp = 1.0 - cos(w);
        if (fp < 1e-300) break;
        const double step = f / fp;
        w -= step;
        if (fabs(step) < 1e-15 * fmax(1.0, w)) break;
    }
    w = fmin(fmax(w, 0.0), 3.141592653589793);
    double s, c;
    sincos(0.5 * w, &s, &c);
    const double inv = s / h;
    q[0] = c, q[1] = hx * inv, q[2] = hy * inv, q[3] = hz * inv;
}

__device__ __forceinline__ bool so3_point(const So3Params &p, long long idx, double (&q)[4]) {
    const int n = 2 * p.n_steps;
    const int k = (int)(idx % n), j = (int)((idx / n) % n), i = (int)(idx / ((long long)n * n));
    const double delta = 2.1450293971110256 / n;  // pi^(2/3) / (2 N)
    const double x = (i - p.n_steps + 0.5) * delta, y = (j - p.n_steps + 0.5) * delta, z = (k - p.n_steps + 0.5) * delta;
    double hx, hy, hz;
    cubochoric_to_homochoric(x, y, z, hx, hy, hz);
    homochoric_to_quat(hx, hy, hz, q);
    if (p.mode == 1) {
        for (int s = 0; s < p.n_sym; ++s) {
            const double w = p.sym[s][0] * q[0] - p.sym[s][1] * q[1] - p.sym[s][2] * q[2] - p.sym[s][3] * q[3];
            if (fabs(w) > q[0] + p.tol) return false;
        }
    } else if (p.mode == 2) {
        if (2.0 * acos(fmin(q[0], 1.0)) > p.max_angle) return false;
    }
    return true;
}

__global__ void __launch_bounds__(SO3_THREADS)
so3_grid_kernel(const So3Params p, const int pass, int *__restrict__ block_counts, const long long *__restrict__ block_offsets,
                double *__restrict__ euler, double *__restrict__ quat) {
    __shared__ int s_warp[SO3_THREADS / 32];
    const long long idx = (long long)blockIdx.x * SO3_THREADS + threadIdx.x;
    double q[4] = {1, 0, 0, 0};
    const bool keep = idx < p.n_points && so3_point(p, idx, q);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SO3_THREADS / 32; ++w) {
        if (w < warp) before += s_warp[w];
        total += s_warp[w];
    }
    if (pass == 0) {
        if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
        return;
    }
    if (!keep) return;
    const long long o = block_offsets[blockIdx.x] + before + __popc(mask & ((1u << lane) - 1u));
    if (p.has_centre) {  // q <- centre * q (Hamilton product)
        const double *c = p.centre;
        const double a = c[0] * q[0] - c[1] * q[1] - c[2] * q[2] - c[3] * q[3];
        const double b = c[0] * q[1] + c[1] * q[0] + c[2] * q[3] - c[3] * q[2];
        const double cc = c[0] * q[2] - c[1] * q[3] + c[2] * q[0] + c[3] * q[1];
        const double d = c[0] * q[3] + c[1] * q[2] - c[2] * q[1] + c[3] * q[0];
        q[0] = a, q[1] = b, q[2] = cc, q[3] = d;
        if (q[0] < 0) q[0] = -q[0], q[1] = -q[1], q[2] = -q[2], q[3] = -q[3];
    }
    if (quat) {  // the active quaternion: the conjugate
        quat[4 * o + 0] = q[0];
        quat[4 * o + 1] = -q[1];
        quat[4 * o + 2] = -q[2];
        quat[4 * o + 3] = -q[3];
    }
    if (euler) {
        // Bunge angles (phi1, Phi, phi2) of the passive matrix of q, exactly as crystal.Rotation.to_euler computes them
        // (the inverse of Rotation.from_euler; orix's convention)
        const double PI = 3.141592653589793;
        const double a = q[0], b = q[1], c = q[2], d = q[3];
        const double om00 = a * a + b * b - c * c - d * d, om22 = a * a - b * b - c * c + d * d;
        const double om01 = 2 * (b * c - a * d), om02 = 2 * (b * d + a * c), om20 = 2 * (b * d - a * c);
        const double om12 = 2 * (c * d - a * b), om21 = 2 * (c * d + a * b);
        double e1, e3;
        const double e2 = acos(fmin(fmax(om22, -1.0), 1.0));
        if (fabs(fabs(om22) - 1.0) <= 1e-8 + 1e-5) {  // numpy.isclose(|om22|, 1)
            e1 = atan2(om01, om00);
            e3 = 0.0;
        } else {
            e1 = atan2(om20, -om21);
            e3 = atan2(om02, om12);
        }
        if (e1 < 0) e1 += 2 * PI;
        if (e3 < 0) e3 += 2 * PI;
        euler[3 * o + 0] = e1 * (180.0 / PI);
        euler[3 * o + 1] = e2 * (180.0 / PI);
        euler[3 * o + 2] = e3 * (180.0 / PI);
    }
}

}  // namespace ds

extern "C" int64_t ds_so3_grid_num_blocks(int32_t n_steps) {
    const long long n = 8ll * n_steps * n_steps * n_steps;
    return (n + ds::SO3_THREADS - 1) / ds::SO3_THREADS;
}

extern "C" int ds_so3_grid(void *stream, int32_t pass, int32_t n_steps, int32_t mode, int32_t n_sym,
                           const double *sym_quats_host, double max_angle_rad, const double *centre_quat_host,
                           int32_t *block_counts, const int64_t *block_offsets, double *euler_deg, double *quat_active) {
    using namespace ds;
    DS_REQUIRE(n_steps > 0 && n_steps <= 1024, "ds_so3_grid: n_steps out of range");
    DS_REQUIRE(mode >= 0 && mode <= 2, "ds_so3_grid: unknown crop mode %d", mode);
    DS_REQUIRE(pass == 0 || pass == 1, "ds_so3_grid: pass must be 0 (count) or 1 (fill)");
    DS_REQUIRE(mode != 1 || (sym_quats_host != nullptr && n_sym >= 1 && n_sym <= SO3_MAX_SYM),
               "ds_so3_grid: the fundamental-zone crop needs 1..24 symmetry quaternions");
    DS_REQUIRE(block_counts != nullptr && (pass == 0 || block_offsets != nullptr), "ds_so3_grid: null block arrays");
    So3Params p;
    p.n_steps = n_steps;
    p.n_points = 8ll * n_steps * n_steps * n_steps;
    p.mode = mode;
    p.n_sym = mode == 1 ? n_sym : 0;
    for (int s = 0; s < p.n_sym; ++s)
        for (int k = 0; k < 4; ++k) p.sym[s][k] = sym_quats_host[4 * s + k];
    p.tol = 1e-9;
    p.max_angle = max_angle_rad;
    p.has_centre = centre_quat_host != nullptr;
    for (int k = 0; k < 4; ++k) p.centre[k] = centre_quat_host ? centre_quat_host[k] : (k == 0 ? 1.0 : 0.0);
    const long long blocks = (p.n_points + SO3_THREADS - 1) / SO3_THREADS;
    so3_grid_kernel<<<(unsigned)blocks, SO3_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        p, pass, block_counts, reinterpret_cast<const long long *>(block_offsets), euler_deg, quat_active);
    return check_launch("ds_so3_grid");
}
