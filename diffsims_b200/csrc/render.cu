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
#include "common.cuh"

namespace ds {

constexpr int RN_WARPS = 8;
constexpr int RN_THREADS = RN_WARPS * 32;
constexpr int RN_RW = 64;  // warp region width  (8 lanes x 8 px)
constexpr int RN_RH = 32;  // warp region height (4 lanes x 8 px)

struct RenderParams {
    int n_tmpl, cap, H, W;
    const int *count;
    const double *xyz;
    const double *intensity;
    double inv_cal_unused, cal, cx, cy, ca, sa, mirror;  // mirror = +1 / -1
    double sigma, clip;
    int radius;  // fast: taps |k| <= radius
    int normalize;
    int table_size;  // fast: hash slots (power of two)
    int n4;          // fast: float4 entries per shifted LUT copy
    float *images;
};

// ---------------------------------------------------------------------------------------------------
// shared-memory carve-up helpers
// ---------------------------------------------------------------------------------------------------
struct FastSmem {
    float4 *lut;               // [4][n4]
    unsigned long long *hash;  // [table_size]
    int *key;                  // [cap]   pixel key of spot j or -1
    float *inten;              // [cap]
    short *ix, *iy;            // [cap]   compacted live spots
    float *amp;                // [cap]
};

__device__ __forceinline__ FastSmem carve_fast(unsigned char *base, const RenderParams &p) {
    FastSmem s;
    s.lut = reinterpret_cast<float4 *>(base);
    base += (size_t)4 * p.n4 * sizeof(float4);
    s.hash = reinterpret_cast<unsigned long long *>(base);
    base += (size_t)p.table_size * 8;
    s.key = reinterpret_cast<int *>(base);
    base += (size_t)p.cap * 4;
    s.inten = reinterpret_cast<float *>(base);
    base += (size_t)p.cap * 4;
    s.amp = reinterpret_cast<float *>(base);
    base += (size_t)p.cap * 4;
    s.ix = reinterpret_cast<short *>(base);
    base += (size_t)p.cap * 2;
    s.iy = reinterpret_cast<short *>(base);
    return s;
}

static size_t fast_smem_bytes(int cap, int n4, int table_size) {
    return (size_t)4 * n4 * 16 + (size_t)table_size * 8 + (size_t)cap * 16;
}

// LUT fetch: 4 consecutive taps L[d .. d+3] of the zero-padded symmetric kernel, any integer offset d.
__device__ __forceinline__ float4 fetch4(const float4 *lut, int n4, int radius, int d) {
    d = max(-(radius + 4), min(d, radius + 1));  // outside the support every tap is 0
    const int a = d + radius + 4;
    return lut[(a & 3) * n4 + (a >> 2)];
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Folded weights of 4 consecutive pixels p0..p0+3 for a delta at pixel `c` on an axis of length n.
__device__ __forceinline__ float4 folded4(const float4 *lut, int n4, int radius, int p0, int c, int n) {
    float4 w = fetch4(lut, n4, radius, p0 - c);
    if (c < radius) w = add4(w, fetch4(lut, n4, radius, p0 + c + 1));               // left/top mirror image
    if (c >= n - radius) w = add4(w, fetch4(lut, n4, radius, p0 + c + 1 - 2 * n));  // right/bottom mirror
    if (radius >= n) {  // kernel wider than the axis: further images at c + 2 n m and -c - 1 + 2 n m
        const int M = radius / n + 1;
        for (int m = -M; m <= M; ++m) {
            if (m != 0) w = add4(w, fetch4(lut, n4, radius, p0 - (c + 2 * n * m)));
            if (m != 0 && m != 1) w = add4(w, fetch4(lut, n4, radius, p0 - (-c - 1 + 2 * n * m)));
        }
    }
    return w;
}

__device__ __forceinline__ void fma_tile(float (&acc)[8][8], const float (&wy)[8], const float (&wx)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wy[i], wx[j], acc[i][j]);
}

// One warp region (64 x 32 px at rx0, ry0): accumulate all live spots into the lane's 8 x 8 tile.
__device__ __forceinline__ void accumulate_fast(const RenderParams &p, const FastSmem &s, int n_live, int rx0,
                                                int ry0, int lane, float (&acc)[8][8]) {
    const int lx = lane & 7, ly = lane >> 3;
    const int x0 = rx0 + 4 * lx, y0 = ry0 + 8 * ly;
    const int R = p.radius;
    const bool wide = (R >= p.W) || (R >= p.H);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int base = 0; base < n_live; base += 32) {
        const int j = base + lane;
        bool hit = false;
        if (j < n_live) {
            const int sx = s.ix[j], sy = s.iy[j];
            hit = wide || (sx + R >= rx0 && sx - R < rx0 + RN_RW && sy + R >= ry0 && sy - R < ry0 + RN_RH);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const int sx = s.ix[base + b], sy = s.iy[base + b];
            const float a = s.amp[base + b];
            const float4 xa = folded4(s.lut, p.n4, R, x0, sx, p.W);
            const float4 xb = folded4(s.lut, p.n4, R, x0 + 32, sx, p.W);
            const float4 ya = folded4(s.lut, p.n4, R, y0, sy, p.H);
            const float4 yb = folded4(s.lut, p.n4, R, y0 + 4, sy, p.H);
            const float wx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float wy[8] = {a * ya.x, a * ya.y, a * ya.z, a * ya.w, a * yb.x, a * yb.y, a * yb.z, a * yb.w};
            fma_tile(acc, wy, wx);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// slow (sub-pixel) path data
// ---------------------------------------------------------------------------------------------------
struct SlowSmem {
    float *fx, *fy, *amp;            // [cap]
    short *xlo, *xhi, *ylo, *yhi;    // [cap]  inclusive clip box
};
__device__ __forceinline__ SlowSmem carve_slow(unsigned char *base, const RenderParams &p) {
    SlowSmem s;
    s.fx = reinterpret_cast<float *>(base);
    s.fy = s.fx + p.cap;
    s.amp = s.fy + p.cap;
    s.xlo = reinterpret_cast<short *>(s.amp + p.cap);
    s.xhi = s.xlo + p.cap;
    s.ylo = s.xhi + p.cap;
    s.yhi = s.ylo + p.cap;
    return s;
}
static size_t slow_smem_bytes(int cap) { return (size_t)cap * 20; }

__device__ __forceinline__ void accumulate_slow(const RenderParams &p, const SlowSmem &s, int n_live, int rx0,
                                                int ry0, int lane, float (&acc)[8][8]) {
    const int lx = lane & 7, ly = lane >> 3;
    const int x0 = rx0 + 4 * lx, y0 = ry0 + 8 * ly;
    const float ef = (float)(-1.0 / (2.0 * p.sigma * p.sigma));
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int base = 0; base < n_live; base += 32) {
        const int j = base + lane;
        bool hit = false;
        if (j < n_live)
            hit = s.xhi[j] >= rx0 && s.xlo[j] < rx0 + RN_RW && s.yhi[j] >= ry0 && s.ylo[j] < ry0 + RN_RH;
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const int k = base + b;
            const float fx = s.fx[k], fy = s.fy[k], a = s.amp[k];
            const int xlo = s.xlo[k], xhi = s.xhi[k], ylo = s.ylo[k], yhi = s.yhi[k];
            float wx[8], wy[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int x = x0 + (q & 3) + (q >> 2) * 32;
                const float dx = (float)x - fx;
                wx[q] = (x >= xlo && x <= xhi) ? __expf(ef * dx * dx) : 0.f;
                const int y = y0 + q;
                const float dy = (float)y - fy;
                wy[q] = (y >= ylo && y <= yhi) ? a * __expf(ef * dy * dy) : 0.f;
            }
            fma_tile(acc, wy, wx);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(RN_THREADS, 2) render_kernel(const RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_n_live, s_n_inframe;
    __shared__ float s_wmax[RN_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x;

    FastSmem fs;
    SlowSmem ss;
    if (FAST) {
        fs = carve_fast(smem_raw, p);
        // normalised 1-D taps, scipy.ndimage._gaussian_kernel1d: exp(-0.5 k^2 / sigma^2) / sum
        __shared__ double s_norm;
        if (warp == 0) {
            double part = 0.0;
            for (int k = lane; k <= p.radius; k += 32)
                part += (k == 0 ? 1.0 : 2.0) * exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_norm = part;
        }
        __syncthreads();
        const double inv = 1.0 / s_norm;
        float *lutf = reinterpret_cast<float *>(fs.lut);
        for (int e = threadIdx.x; e < 16 * p.n4; e += RN_THREADS) {
            const int copy = e / (4 * p.n4), rem = e % (4 * p.n4);
            const int a = (rem >> 2) * 4 + copy + (rem & 3);  // position in the padded kernel
            const int k = abs(a - (p.radius + 4));
            lutf[e] = (k <= p.radius) ? (float)(exp(-0.5 / (p.sigma * p.sigma) * (double)k * (double)k) * inv) : 0.f;
        }
        for (int e = threadIdx.x; e < p.table_size; e += RN_THREADS) fs.hash[e] = 0ull;
    } else {
        ss = carve_slow(smem_raw, p);
    }
    if (threadIdx.x == 0) s_n_live = s_n_inframe = 0;
    __syncthreads();

    // ---- project spots to detector pixels (simulation2d.py:261-285, :422-430), float64 ---------------
    const int n = min(p.count[t], p.cap);
    const size_t row = (size_t)t * p.cap;
    if (FAST) {
        for (int j = threadIdx.x; j < n; j += RN_THREADS) {
            const double xs = p.xyz[3 * (row + j)] / p.cal, ys = p.xyz[3 * (row + j) + 1] / p.cal;
            // r cos(+-atan2(y, x) + a) + cx written without the polar round trip
            const double px = xs * p.ca - p.mirror * ys * p.sa + p.cx;
            const double py = p.mirror * ys * p.ca + xs * p.sa + p.cy;
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
            fs.inten[j] = (float)p.intensity[row + j];
        }
        __syncthreads();
        if (warp == 0) {  // deterministic compaction of the surviving spots
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
                    fs.ix[d] = (short)(key % p.W);
                    fs.iy[d] = (short)(key / p.W);
                    fs.amp[d] = fs.inten[j];
                }
                n_live += __popc(mask);
            }
            if (lane == 0) s_n_live = n_live;
        }
    } else {
        if (warp == 0) {
            const double pref = 1.0 / (2.0 * 3.141592653589793 * p.sigma * p.sigma);
            const double ef = -1.0 / (2.0 * p.sigma * p.sigma);
            int n_live = 0, n_in = 0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                bool live = false, inframe = false;
                double px = 0, py = 0, I = 0, rad = 0;
                if (j < n) {
                    const double xs = p.xyz[3 * (row + j)] / p.cal, ys = p.xyz[3 * (row + j) + 1] / p.cal;
                    px = xs * p.ca - p.mirror * ys * p.sa + p.cx;
                    py = p.mirror * ys * p.ca + xs * p.sa + p.cy;
                    I = p.intensity[row + j];
                    if (px >= 0.0 && px < (double)p.W && py >= 0.0 && py < (double)p.H) {
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
                    ss.xlo[d] = (short)max(0, (int)fmin(ceil(px - rad), 32000.0));
                    ss.xhi[d] = (short)(min(p.W, (int)fmin(floor(px + rad + 1.0), 32000.0)) - 1);
                    ss.ylo[d] = (short)max(0, (int)fmin(ceil(py - rad), 32000.0));
                    ss.yhi[d] = (short)(min(p.H, (int)fmin(floor(py + rad + 1.0), 32000.0)) - 1);
                }
                n_live += __popc(mask);
            }
            if (lane == 0) {
                s_n_live = n_live;
                s_n_inframe = n_in;
            }
        }
    }
    __syncthreads();
    const int n_live = s_n_live;

    const int nrx = (p.W + RN_RW - 1) / RN_RW, nry = (p.H + RN_RH - 1) / RN_RH;
    const int n_regions = nrx * nry;
    float *img = p.images + (size_t)t * p.H * p.W;
    const int lx = lane & 7, ly = lane >> 3;
    const bool vec_ok = (p.W & 3) == 0;

    float scale = 1.f;
    // reference: no spot in frame -> zeros, returned before the normalisation (simulation2d.py:434-435)
    // (the test is on the IN-FRAME spots; slow-path spots skipped for a NaN radius still count, so an
    // all-skipped pattern is 0 / 0 = NaN in the reference and here)
    const int n_pass = (p.normalize && (FAST ? n_live > 0 : s_n_inframe > 0)) ? 2 : 1;
    for (int pass = 0; pass < n_pass; ++pass) {
        const bool store = (pass == n_pass - 1);
        float vmax = -INFINITY;
        for (int reg = warp; reg < n_regions; reg += RN_WARPS) {
            const int rx0 = (reg % nrx) * RN_RW, ry0 = (reg / nrx) * RN_RH;
            float acc[8][8];
            if (FAST)
                accumulate_fast(p, fs, n_live, rx0, ry0, lane, acc);
            else
                accumulate_slow(p, ss, n_live, rx0, ry0, lane, acc);
            const int x0 = rx0 + 4 * lx, y0 = ry0 + 8 * ly;
            if (!store) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int x = x0 + (j & 3) + (j >> 2) * 32, y = y0 + i;
                        if (x < p.W && y < p.H) vmax = fmaxf(vmax, acc[i][j]);
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int y = y0 + i;
                    if (y >= p.H) continue;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int x = x0 + 32 * h;
                        float4 v = make_float4(acc[i][4 * h] * scale, acc[i][4 * h + 1] * scale,
                                               acc[i][4 * h + 2] * scale, acc[i][4 * h + 3] * scale);
                        float *dst = img + (size_t)y * p.W + x;
                        if (vec_ok && x + 3 < p.W) {
                            __stcs(reinterpret_cast<float4 *>(dst), v);
                        } else {
                            if (x < p.W) dst[0] = v.x;
                            if (x + 1 < p.W) dst[1] = v.y;
                            if (x + 2 < p.W) dst[2] = v.z;
                            if (x + 3 < p.W) dst[3] = v.w;
                        }
                    }
                }
            }
        }
        if (!store) {
            vmax = warp_max(vmax);
            if (lane == 0) s_wmax[warp] = vmax;
            __syncthreads();
            float m = s_wmax[0];
#pragma unroll
            for (int k = 1; k < RN_WARPS; ++k) m = fmaxf(m, s_wmax[k]);
            scale = 1.f / m;  // np.divide(pattern, np.max(pattern)), simulation2d.py:440-441
        }
    }
}

}  // namespace ds

extern "C" int ds_render(void *stream, int32_t n_tmpl, int32_t cap, const int32_t *count, const double *xyz,
                         const double *intensity, int32_t H, int32_t W, double calibration, double cx, double cy,
                         double in_plane_angle_deg, int32_t mirrored, int32_t fast, double sigma, int32_t radius,
                         double clip_threshold, int32_t normalize, float *images) {
    using namespace ds;
    DS_REQUIRE(n_tmpl >= 0 && cap > 0 && H > 0 && W > 0, "ds_render: bad sizes");
    DS_REQUIRE(H <= 16384 && W <= 16384, "ds_render: image larger than 16384 px per side");
    DS_REQUIRE(calibration != 0.0, "ds_render: calibration cannot be zero");
    DS_REQUIRE(sigma > 0.0, "ds_render: sigma must be positive");
    DS_REQUIRE((reinterpret_cast<uintptr_t>(images) & 15) == 0, "ds_render: images must be 16-byte aligned");
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
    p.inv_cal_unused = 0.0;
    p.cx = cx;
    p.cy = cy;
    const double ang = in_plane_angle_deg * (3.141592653589793 / 180.0);
    p.ca = cos(ang);
    p.sa = sin(ang);
    if (in_plane_angle_deg == 0.0) {
        p.ca = 1.0;
        p.sa = 0.0;
    }
    p.mirror = mirrored ? -1.0 : 1.0;
    p.sigma = sigma;
    p.clip = clip_threshold;
    p.radius = radius;
    p.normalize = normalize;
    p.images = images;
    p.table_size = 0;
    p.n4 = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // vector stores need every row start 16-byte aligned
    if ((W & 3) != 0) { /* scalar path inside the kernel */
    }
    if (fast) {
        DS_REQUIRE(radius >= 0 && radius <= 4096, "ds_render: radius out of range");
        int ts = 64;
        while (ts < 2 * cap) ts <<= 1;
        p.table_size = ts;
        p.n4 = ((2 * radius + 5) >> 2) + 2;
        const size_t smem = fast_smem_bytes(cap, p.n4, ts);
        DS_REQUIRE(smem <= 200 * 1024, "ds_render: cap=%d / radius=%d need %zu bytes of shared memory", cap, radius,
                   smem);
        static size_t attr = 0;
        if (smem > 48 * 1024 && smem > attr) {
            cudaFuncSetAttribute(render_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr = smem;
        }
        render_kernel<true><<<n_tmpl, RN_THREADS, smem, st>>>(p);
    } else {
        const size_t smem = slow_smem_bytes(cap);
        DS_REQUIRE(smem <= 200 * 1024, "ds_render: cap=%d needs %zu bytes of shared memory", cap, smem);
        static size_t attr = 0;
        if (smem > 48 * 1024 && smem > attr) {
            cudaFuncSetAttribute(render_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr = smem;
        }
        render_kernel<false><<<n_tmpl, RN_THREADS, smem, st>>>(p);
    }
    return check_launch("ds_render");
}
