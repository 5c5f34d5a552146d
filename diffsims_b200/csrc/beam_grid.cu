// Rotation-list producer on the device: beam directions inside the stereographic
// triangle of a crystal system, as Bunge Euler angles with phi1 = 0.
//
// Replaces get_beam_directions_grid (diffsims/generators/rotation_list_generators.py:176-267) for the cube
// meshes of get_cube_mesh_vertices (diffsims/generators/sphere_mesh_generators.py:96-197) followed by
// beam_directions_grid_to_euler (:486-526), so that a 3e5 - 1e6 entry grid is produced directly in HBM
// (optionally as the active quaternions the simulate kernel consumes) instead of as a Python list.
// The 1-D face grid i[k] (tan-spaced, depends on the mesh type) is computed by the host exactly as the
// reference does; the kernel forms the 6 n^2 + 2 cube points in the reference's order (bottom, top, east,
// west, south, north, two corners), normalises, crops with the three plane tests, and compacts in order:
// pass 0 counts survivors per block, pass 1 writes them at the scanned block offsets.
// The uv-sphere, icosahedral and random meshes (:42-93, :378-483) are small host-side vertex lists; ds_beam_points
// runs the same crop / compaction / conversion on vertices already in device memory.
#include "common.cuh"

namespace ds {

struct BeamGridParams {
    int n_i;
    long long n_points;
    const double *i_vals;
    const double *points;  // non-null: [n_points][3] mesh vertices instead of the cube faces
    int mode;  // 0 no crop (triclinic) | 1 x >= eps (monoclinic, as the reference computes it) | 2 triangle
    double nrm[9];
    double eps;
};

// crop to the stereographic triangle, rotation_list_generators.py:237-263
__device__ __forceinline__ bool beam_crop(const BeamGridParams &p, double vx, double vy, double vz) {
    if (p.mode == 0) return true;
    if (p.mode == 1) return vx >= p.eps;
    return (p.nrm[0] * vx + p.nrm[1] * vy + p.nrm[2] * vz >= p.eps) &&
           (p.nrm[3] * vx + p.nrm[4] * vy + p.nrm[5] * vz >= p.eps) &&
           (p.nrm[6] * vx + p.nrm[7] * vy + p.nrm[8] * vz >= p.eps);
}

__device__ __forceinline__ bool beam_point(const BeamGridParams &p, long long idx, double &vx, double &vy,
                                           double &vz) {
    if (p.points) {
        vx = p.points[3 * idx], vy = p.points[3 * idx + 1], vz = p.points[3 * idx + 2];
        return beam_crop(p, vx, vy, vz);
    }
    const long long nn = (long long)p.n_i * p.n_i;
    const int face = (int)(idx / nn);
    double x, y, z = 1.0;
    if (face < 6) {
        const long long rem = idx - (long long)face * nn;
        x = p.i_vals[rem % p.n_i];  // np.meshgrid(i, i): x varies fastest
        y = p.i_vals[rem / p.n_i];
    } else {
        x = y = 0;
    }
    switch (face) {
        case 0: vx = -x, vy = -y, vz = -z; break;  // bottom
        case 1: vx = x, vy = y, vz = z; break;     // top
        case 2: vx = z, vy = x, vz = -y; break;    // east
        case 3: vx = -z, vy = -x, vz = y; break;   // west
        case 4: vx = x, vy = -z, vz = y; break;    // south
        case 5: vx = -x, vy = z, vz = -y; break;   // north
        default:                                   // the two corners the faces miss
            if (idx - 6 * nn == 0) vx = -1, vy = 1, vz = 1; else vx = 1, vy = -1, vz = -1;
    }
    const double inv = sqrt(vx * vx + vy * vy + vz * vz);
    vx /= inv, vy /= inv, vz /= inv;
    return beam_crop(p, vx, vy, vz);
}

constexpr int BG_THREADS = 256;

__global__ void __launch_bounds__(BG_THREADS)
beam_grid_kernel(const BeamGridParams p, const int pass, int *__restrict__ block_counts,
                 const long long *__restrict__ block_offsets, double *__restrict__ euler,
                 double *__restrict__ quat) {
    __shared__ int s_warp[BG_THREADS / 32];
    const long long idx = (long long)blockIdx.x * BG_THREADS + threadIdx.x;
    double vx = 0, vy = 0, vz = 0;
    const bool keep = idx < p.n_points && beam_point(p, idx, vx, vy, vz);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BG_THREADS / 32; ++w) {
        if (w < warp) before += s_warp[w];
        total += s_warp[w];
    }
    if (pass == 0) {
        if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
        return;
    }
    if (!keep) return;
    const long long o = block_offsets[blockIdx.x] + before + __popc(mask & ((1u << lane) - 1u));
    // beam_directions_grid_to_euler, sphere_mesh_generators.py:508-526
    const double PI = 3.141592653589793;
    const double norm = sqrt(vx * vx + vy * vy + vz * vz);
    const double Phi = acos(vz / norm);
    const double norm_proj = sqrt(vx * vx + vy * vy);
    double sign = (vy > 0.0) ? 1.0 : ((vy < 0.0) ? -1.0 : 0.0);
    if (vy == 0.0) sign = (vx > 0.0) ? 1.0 : ((vx < 0.0) ? -1.0 : 0.0);
    double ac = acos(vx / norm_proj);
    if (isnan(ac)) ac = 0.0;  // np.nan_to_num
    const double phi2 = PI / 2 - sign * ac;
    if (euler) {
        euler[3 * o + 0] = 0.0;
        euler[3 * o + 1] = Phi * (180.0 / PI);
        euler[3 * o + 2] = phi2 * (180.0 / PI);
    }
    if (quat) {
        // orix Rotation.from_euler((0, Phi, phi2)) then inverted: the ACTIVE quaternion K2 applies
        double sh, ch, ss, cs;
        sincos(0.5 * Phi, &sh, &ch);
        sincos(0.5 * phi2, &ss, &cs);
        double a = ch * cs, b = -sh * cs, c = sh * ss, d = -ch * ss;  // sigma = phi2/2, delta = -phi2/2
        if (a < 0) a = -a, b = -b, c = -c, d = -d;
        quat[4 * o + 0] = a;
        quat[4 * o + 1] = -b;
        quat[4 * o + 2] = -c;
        quat[4 * o + 3] = -d;
    }
}

}  // namespace ds

extern "C" int64_t ds_beam_grid_num_blocks(int32_t n_i) {
    const long long n = 6ll * n_i * n_i + 2;
    return (n + ds::BG_THREADS - 1) / ds::BG_THREADS;
}

static int beam_launch(void *stream, const char *what, ds::BeamGridParams &p, int32_t pass, int32_t mode,
                       const double *normals_host, double epsilon, int32_t *block_counts,
                       const int64_t *block_offsets, double *euler_deg, double *quat_active) {
    using namespace ds;
    DS_REQUIRE(mode >= 0 && mode <= 2, "%s: unknown crop mode %d", what, mode);
    DS_REQUIRE(pass == 0 || pass == 1, "%s: pass must be 0 (count) or 1 (fill)", what);
    DS_REQUIRE(mode != 2 || normals_host != nullptr, "%s: triangle crop needs three plane normals", what);
    DS_REQUIRE(block_counts != nullptr && (pass == 0 || block_offsets != nullptr), "%s: null block arrays", what);
    p.mode = mode;
    for (int k = 0; k < 9; ++k) p.nrm[k] = normals_host ? normals_host[k] : 0.0;
    p.eps = epsilon;
    const long long blocks = (p.n_points + BG_THREADS - 1) / BG_THREADS;
    beam_grid_kernel<<<(unsigned)blocks, BG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        p, pass, block_counts, reinterpret_cast<const long long *>(block_offsets), euler_deg, quat_active);
    return check_launch(what);
}

extern "C" int ds_beam_grid(void *stream, int32_t pass, int32_t n_i, const double *i_vals, int32_t mode,
                            const double *normals_host, double epsilon, int32_t *block_counts,
                            const int64_t *block_offsets, double *euler_deg, double *quat_active) {
    using namespace ds;
    DS_REQUIRE(n_i > 0 && n_i <= 32768, "ds_beam_grid: n_i out of range");
    DS_REQUIRE(i_vals != nullptr, "ds_beam_grid: null face grid");
    BeamGridParams p;
    p.n_i = n_i;
    p.n_points = 6ll * n_i * n_i + 2;
    p.i_vals = i_vals;
    p.points = nullptr;
    return beam_launch(stream, "ds_beam_grid", p, pass, mode, normals_host, epsilon, block_counts, block_offsets,
                       euler_deg, quat_active);
}

extern "C" int64_t ds_beam_points_num_blocks(int64_t n_points) {
    return (n_points + ds::BG_THREADS - 1) / ds::BG_THREADS;
}

extern "C" int ds_beam_points(void *stream, int32_t pass, int64_t n_points, const double *points, int32_t mode,
                              const double *normals_host, double epsilon, int32_t *block_counts,
                              const int64_t *block_offsets, double *euler_deg, double *quat_active) {
    using namespace ds;
    DS_REQUIRE(n_points > 0 && n_points <= (1ll << 40), "ds_beam_points: n_points out of range");
    DS_REQUIRE(points != nullptr, "ds_beam_points: null vertex array");
    BeamGridParams p;
    p.n_i = 0;
    p.n_points = n_points;
    p.i_vals = nullptr;
    p.points = points;
    return beam_launch(stream, "ds_beam_points", p, pass, mode, normals_host, epsilon, block_counts, block_offsets,
                       euler_deg, quat_active);
}
