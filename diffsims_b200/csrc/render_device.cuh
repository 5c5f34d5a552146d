// Device helpers shared by the render kernels (render.cu, render_pipe.cu): parameter block, shared-memory
// carve-up, the shifted tap LUT, reflect folding, the per-region accumulate functions.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace ds {

int num_sms();

constexpr int RN_WARPS = 8;
constexpr int RN_THREADS = RN_WARPS * 32;
constexpr int RN_RW = 64;  // warp region width  (8 lanes x 8 px)
constexpr int RN_RH = 32;  // warp region height (4 lanes x 8 px)
// Zero padding of the tap LUT on either side of the kernel support.  A spot that passes the region cull is at most
// R + 31 pixels from any 4-pixel group of the region, so with 32 the hot loop needs no index clamp.
constexpr int LUT_PAD = 32;
inline int lut_entries(int radius) { return ((2 * radius + 2 * LUT_PAD + 3) >> 2) + 2; }

struct RenderParams {
    int n_tmpl, cap, H, W;
    const int *count;
    const double *xyz;
    const double *intensity;
    double cal, cx, cy, ca, sa, mirror;  // mirror = +1 / -1
    double sigma, clip;
    int radius;  // fast: taps |k| <= radius
    int normalize;
    int table_size;  // fast: hash slots (power of two)
    int n4;          // fast: float4 entries per shifted LUT copy
    int stage;       // 1: spot rows are prefetched into shared memory with cp.async.bulk (TMA engine)
    int keep_outside;  // sub-pixel path: do not drop spots whose centre lies outside the frame
    float *images;
    int *ticket;  // [2] device scratch, zero on entry and on exit: dynamic template assignment
};

// ---------------------------------------------------------------------------------------------------
// shared-memory carve-up helpers
// ---------------------------------------------------------------------------------------------------
struct FastSmem {
    float4 *lut;               // [4][n4]
    unsigned long long *hash;  // [table_size]
    int *key;                  // [cap]   pixel key of spot j or -1
    float *inten;              // [cap]
    uint2 *spot;               // [cap]   compacted live spots: .x = ix | (iy | fold << 14) << 16, .y = amplitude bits
};

__device__ __forceinline__ int spot_ix(uint2 r) { return (int)(r.x & 0xffffu); }
__device__ __forceinline__ int spot_iy(uint2 r) { return (int)((r.x >> 16) & 0x3fffu); }
__device__ __forceinline__ bool spot_fold(uint2 r) { return (r.x >> 30) & 1u; }
__device__ __forceinline__ float spot_amp(uint2 r) { return __uint_as_float(r.y); }

__device__ __forceinline__ FastSmem carve_fast(unsigned char *base, const RenderParams &p) {
    FastSmem s;
    s.lut = nullptr;  // CTA-wide, set by the kernel
    s.hash = reinterpret_cast<unsigned long long *>(base);
    base += (size_t)p.table_size * 8;
    s.key = reinterpret_cast<int *>(base);
    base += (size_t)p.cap * 4;
    s.inten = reinterpret_cast<float *>(base);
    base += (size_t)p.cap * 4;
    s.spot = reinterpret_cast<uint2 *>(base);
    return s;
}


// LUT layout.  The padded symmetric kernel is L[a] = w[|a - (R + LUT_PAD)|] (0 outside the support).  Copy k
// (k = 0..3) holds float4 entries i -> (L[4i+k] .. L[4i+k+3]), so four consecutive taps at ANY integer
// offset are one aligned LDS.128: offset d -> a = d + R + LUT_PAD, copy a & 3, entry a >> 2 clamped to
// [0, n4 - 1] (both end entries are all zero).  Lanes of a quarter warp read consecutive entries.
__device__ __forceinline__ float4 fetch4(const float4 *lut, int n4, int R, int d) {
    const int a = d + R + LUT_PAD;
    const int i = min(max(a >> 2, 0), n4 - 1);
    return lut[(a & 3) * n4 + i];
}
__device__ __forceinline__ float tap(const float4 *lut, int n4, int R, int d) {  // w[|d|], 0 outside
    return fetch4(lut, n4, R, d).x;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Folded weights of 4 consecutive pixels p0..p0+3 for a delta at pixel `c` on an axis of length n:
// direct tap + the mirror images of scipy's mode="reflect" (... d c b a | a b c d | d c b a ...).
template <bool WIDE>
__device__ __forceinline__ float4 folded4(const float4 *lut, int n4, int R, int p0, int c, int n) {
    float4 w = fetch4(lut, n4, R, p0 - c);
    if (!WIDE) {
        if (c < R) w = add4(w, fetch4(lut, n4, R, p0 + c + 1));               // image at -c - 1
        if (c >= n - R) w = add4(w, fetch4(lut, n4, R, p0 + c + 1 - 2 * n));  // image at 2n - 1 - c
    } else {  // kernel wider than the axis: images at c + 2 n m and -c - 1 + 2 n m
        const int M = R / n + 1;
        for (int m = -M; m <= M; ++m) {
            if (m != 0) w = add4(w, fetch4(lut, n4, R, p0 - (c + 2 * n * m)));
            w = add4(w, fetch4(lut, n4, R, p0 - (-c - 1 + 2 * n * m)));
        }
    }
    return w;
}

// Upper bound of the folded weight over the pixel interval [lo, hi] (taps decrease with distance).
__device__ __forceinline__ float folded_bound(const float4 *lut, int n4, int R, int lo, int hi, int c, int n) {
    auto dist = [&](int pos) { return pos < lo ? lo - pos : (pos > hi ? pos - hi : 0); };
    float b = tap(lut, n4, R, dist(c));
    if (c < R) b += tap(lut, n4, R, dist(-c - 1));
    if (c >= n - R) b += tap(lut, n4, R, dist(2 * n - 1 - c));
    return b;
}

__device__ __forceinline__ void fma_tile(float (&acc)[8][8], const float (&wy)[8], const float (&wx)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wy[i], wx[j], acc[i][j]);
}

// ---- hot path helpers: 32-bit shared-window addresses and explicit ld.shared (the generic-pointer form
// costs an S2R + LEA address conversion per access) ------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
struct LutRef {
    uint32_t base;  // shared address of copy 0
    int n4, last;   // entries per copy, n4 - 1
    int bias;       // R + LUT_PAD
};
// taps L[a .. a+3] with a = offset + R + LUT_PAD already applied
__device__ __forceinline__ float4 fetch4s(const LutRef &L, int a) {
    const int i = min(max(a >> 2, 0), L.last);
    return lds128(L.base + (uint32_t)(((a & 3) * L.n4 + i) << 4));
}
// shared address of the entry holding taps a .. a + 3, for a known to lie inside the padded table
__device__ __forceinline__ uint32_t fetch_addr(const LutRef &L, int a) {
    return L.base + (uint32_t)(((a & 3) * L.n4 + (a >> 2)) << 4);
}
__device__ __forceinline__ float4 folded4s(const LutRef &L, int p0b /* p0 + R + LUT_PAD */, int c, int n, int R) {
    float4 w = fetch4s(L, p0b - c);
    if (c < R) w = add4(w, fetch4s(L, p0b + c + 1));               // image at -c - 1
    if (c >= n - R) w = add4(w, fetch4s(L, p0b + c + 1 - 2 * n));  // image at 2n - 1 - c
    return w;
}

// One warp region (64 x 32 px at rx0, ry0): accumulate all live spots into the lane's 8 x 8 tile.
// Returns false (acc untouched) when no spot reaches the region.  The two 32-px column groups of the
// region are culled separately (a spot's box is 2R+1 wide, the region 64).
template <bool WIDE>
__device__ __forceinline__ bool accumulate_fast(const RenderParams &p, const FastSmem &s, int n_live, int rx0,
                                                int ry0, int lane, float (&acc)[8][8]) {
    const int lx = lane & 7, ly = lane >> 3;
    const int x0 = rx0 + 4 * lx, y0 = ry0 + 8 * ly;
    const int R = p.radius;
    LutRef L;
    L.base = smem_u32(s.lut);
    L.n4 = p.n4;
    L.last = p.n4 - 1;
    L.bias = R + LUT_PAD;
    const uint32_t spot_s = smem_u32(s.spot);
    const int xb = x0 + L.bias, yb = y0 + L.bias;
    bool any = false;
    for (int base = 0; base < n_live; base += 32) {
        const int j = base + lane;
        bool hit = false;
        if (j < n_live) {
            const uint2 r = lds64(spot_s + 8u * j);
            const int sx = spot_ix(r), sy = spot_iy(r);
            hit = WIDE || (sx + R >= rx0 && sx - R < rx0 + RN_RW && sy + R >= ry0 && sy - R < ry0 + RN_RH);
        }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask && !any) {
            any = true;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
        }
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint2 r = lds64(spot_s + 8u * (base + b));
            const int sx = spot_ix(r), sy = spot_iy(r);
            const float a = spot_amp(r);
            float4 ya, yb4;
            uint32_t ay_s = 0;
            const uint32_t ax_s = fetch_addr(L, xb - sx);  // column group h: + 8 entries
            if (WIDE) {
                ya = folded4<true>(s.lut, p.n4, R, y0, sy, p.H);
                yb4 = folded4<true>(s.lut, p.n4, R, y0 + 4, sy, p.H);
            } else if (spot_fold(r)) {  // the box crosses a border: add the reflect-folded images
                ya = folded4s(L, yb, sy, p.H, R);
                yb4 = folded4s(L, yb + 4, sy, p.H, R);
            } else {  // in range by the cull above: no clamp; rows y0 + 4 .. + 7 are the next entry of the same copy
                ay_s = fetch_addr(L, yb - sy);
                ya = lds128(ay_s);
                yb4 = lds128(ay_s + 16u);
            }
            const float wy[8] = {a * ya.x, a * ya.y, a * ya.z, a * ya.w, a * yb4.x, a * yb4.y, a * yb4.z, a * yb4.w};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // column group h covers x in [rx0 + 32 h, rx0 + 32 h + 31]
                if (!WIDE && !(sx + R >= rx0 + 32 * h && sx - R < rx0 + 32 * h + 32)) continue;
                float4 xw;
                if (WIDE)
                    xw = folded4<true>(s.lut, p.n4, R, x0 + 32 * h, sx, p.W);
                else if (spot_fold(r))
                    xw = folded4s(L, xb + 32 * h, sx, p.W, R);
                else
                    xw = lds128(ax_s + 128u * h);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][4 * h + 0] = fmaf(wy[i], xw.x, acc[i][4 * h + 0]);
                    acc[i][4 * h + 1] = fmaf(wy[i], xw.y, acc[i][4 * h + 1]);
                    acc[i][4 * h + 2] = fmaf(wy[i], xw.z, acc[i][4 * h + 2]);
                    acc[i][4 * h + 3] = fmaf(wy[i], xw.w, acc[i][4 * h + 3]);
                }
            }
        }
    }
    return any;
}

// Stream one finished warp region to the image (acc * sc, or zeros when no spot reached it) with evict-first
// 16-byte stores: per store instruction a quarter warp covers 128 contiguous bytes of one row.  Regions
// entirely inside the frame (all of them when H % 32 == 0 and W % 64 == 0) take a branch-free path.
template <bool VEC>
__device__ __forceinline__ void store_region(const RenderParams &p, float *img, int rx0, int ry0, int lane,
                                             const float (&acc)[8][8], bool any, float sc) {
    const int lx = lane & 7, ly = lane >> 3;
    const int x0 = rx0 + 4 * lx, y0 = ry0 + 8 * ly;
    float *dst = img + (size_t)y0 * p.W + x0;
    if (VEC && ry0 + RN_RH <= p.H && rx0 + RN_RW <= p.W) {
        float4 *d = reinterpret_cast<float4 *>(dst);
        const int row4 = p.W >> 2;
        if (!any) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __stcs(d, z);
                __stcs(d + 8, z);
                d += row4;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __stcs(d, make_float4(acc[i][0] * sc, acc[i][1] * sc, acc[i][2] * sc, acc[i][3] * sc));
                __stcs(d + 8, make_float4(acc[i][4] * sc, acc[i][5] * sc, acc[i][6] * sc, acc[i][7] * sc));
                d += row4;
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool yok = y0 + i < p.H;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float *d = dst + (size_t)i * p.W + 32 * h;
            if (VEC) {  // W % 4 == 0: a float4 group is entirely inside or outside the frame
                if (yok && x0 + 32 * h < p.W)
                    __stcs(reinterpret_cast<float4 *>(d),
                           any ? make_float4(acc[i][4 * h] * sc, acc[i][4 * h + 1] * sc, acc[i][4 * h + 2] * sc,
                                             acc[i][4 * h + 3] * sc)
                               : make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (yok && x0 + 32 * h + q < p.W) d[q] = any ? acc[i][4 * h + q] * sc : 0.f;
            }
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
inline size_t slow_smem_bytes(int cap) { return (size_t)cap * 20; }

__device__ __forceinline__ bool accumulate_slow(const RenderParams &p, const SlowSmem &s, int n_live, int rx0,
                                                int ry0, int lane, float (&acc)[8][8]) {
    const int lx = lane & 7, ly = lane >> 3;
    const int x0 = rx0 + 4 * lx, y0 = ry0 + 8 * ly;
    const float ef = (float)(-1.0 / (2.0 * p.sigma * p.sigma));
    bool any = false;
    for (int base = 0; base < n_live; base += 32) {
        const int j = base + lane;
        bool hit = false;
        if (j < n_live)
            hit = s.xhi[j] >= rx0 && s.xlo[j] < rx0 + RN_RW && s.yhi[j] >= ry0 && s.ylo[j] < ry0 + RN_RH;
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask && !any) {
            any = true;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) acc[i][jj] = 0.f;
        }
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
    return any;
}

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------

}  // namespace ds
