// Device helpers shared by the render kernels (render.cu, render_pipe.cu): parameter block, shared-memory
// carve-up, the shifted tap LUT, reflect folding, the per-region accumulate functions.
#pragma once
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace ds {

int num_sms();
// bytes of one prepared-template record of the tcgen05 render path (render_prep.cu, render_umma.cu)
//   int32 t, n_live, n_half[2], pad[4] | uint2 spot[cap] | uint16 list[2][cap] | uint32 window[2][ceil(cap / 16)]
__host__ __device__ inline int umma_windows_offset(int cap) { return 32 + cap * 8 + 2 * cap * 2; }
__host__ __device__ inline int umma_record_bytes(int cap) { return (umma_windows_offset(cap) + 2 * ((cap + 15) / 16) * 4 + 15) & ~15; }

// record of the row-binned tcgen05 render path (render_rows.cu)
//   int32 t, n_live, uint32 chunk mask, pad[5] | uint2 spot[cap] (column | row << 16, amplitude), ordered by row |
//   uint16 row_offset[H + 1 <= 257, padded]
__host__ __device__ inline int rows_offsets_offset(int cap) { return 32 + cap * 8; }
__host__ __device__ inline int rows_record_bytes(int cap) { return (rows_offsets_offset(cap) + 2 * 264 + 15) & ~15; }

constexpr int RN_WARPS = 8;
constexpr int RN_THREADS = RN_WARPS * 32;
constexpr int RN_RW = 64;  // warp region width  (8 lanes x 8 px)
constexpr int RN_RH = 32;  // warp region height (4 lanes x 8 px)
// Zero padding of the tap LUT on either side of the kernel support.  A spot that passes the region cull is at most
// R + 63 pixels from any pixel of the region (R + 31 within a culled 32-px column group), so with 64 the hot
// loops need no index clamp.
constexpr int LUT_PAD = 64;
inline int lut_entries(int radius) { return ((2 * radius + 2 * LUT_PAD + 3) >> 2) + 2; }
// shared bytes of the tap tables: four shifted float copies (n4 float4 each) + the packed bf16 hi|lo table
__host__ __device__ inline size_t lut_smem_bytes(int n4) { return (size_t)5 * n4 * 16; }

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
    int hits_bytes;    // fast path: bytes of the per-warp hit lists behind the LUT (0 = tensor-core path off)
    int mma_min;       // spots per region from which the tensor-core path is taken
    int mma_tmpl_min;  // spots per template from which regions are examined for it
    float *images;
    int *ticket;  // [2] device scratch, zero on entry and on exit: dynamic template assignment
    double mean_spots_hint;  // caller's estimate of the mean reflections per template (<= 0: unknown)
};

// Detector pixel coordinates of a spot (simulation2d.py:261-285: r cos(+-atan2(y, x) + a) + cx, written without the polar
// round trip).  Every operation is individually rounded IEEE double arithmetic in this fixed order -- no fused
// multiply-add -- so the pixel a coordinate truncates to is exactly what evaluating
//     px = (x / cal) * cos(a) - (m * (y / cal)) * sin(a) + cx,   py = (m * (y / cal)) * cos(a) + (x / cal) * sin(a) + cy
// in numpy gives, on every kernel and every compiler version (knife-edge pixels are decided reproducibly).
__device__ __forceinline__ void project_spot(const RenderParams &p, double x, double y, double &px, double &py) {
    const double xs = __ddiv_rn(x, p.cal), ys = __dmul_rn(p.mirror, __ddiv_rn(y, p.cal));
    px = __dadd_rn(__dsub_rn(__dmul_rn(xs, p.ca), __dmul_rn(ys, p.sa)), p.cx);
    py = __dadd_rn(__dadd_rn(__dmul_rn(ys, p.ca), __dmul_rn(xs, p.sa)), p.cy);
}

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

// Fill the tap tables (all threads of the CTA; `inv_norm` = 1 / sum of the unnormalised taps):
//   floats [0, 16 n4):        the four shifted copies described above (copy 0 is the flat padded kernel L[a]);
//   words  [16 n4, 20 n4):    L[a] split for the tensor-core path, bf16(L) in the low half and
//                             bf16(L - bf16(L)) in the high half.
__device__ __forceinline__ void fill_lut(float4 *lut, int n4, int radius, double sigma, double inv_norm, int tid,
                                         int n_threads) {
    float *lutf = reinterpret_cast<float *>(lut);
    uint32_t *hl = reinterpret_cast<uint32_t *>(lutf + 16 * n4);
    for (int e = tid; e < 20 * n4; e += n_threads) {
        const int copy = e / (4 * n4), rem = e % (4 * n4);
        const int a = (copy < 4) ? (rem >> 2) * 4 + copy + (rem & 3) : rem;  // position in the padded kernel
        const int k = abs(a - (radius + LUT_PAD));
        const float w = (k <= radius) ? (float)(exp(-0.5 / (sigma * sigma) * (double)k * (double)k) * inv_norm) : 0.f;
        if (copy < 4) {
            lutf[e] = w;
        } else {
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            hl[rem] = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
        }
    }
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
// (spot records are rewritten between templates: volatile + memory clobber keeps the load behind the barrier / __syncwarp
// that orders it after the writer; the tap-table loads above read data that never changes after the kernel's prologue)
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
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

// ---------------------------------------------------------------------------------------------------
// Dense patterns: a region reached by S spots is the rank-S product  out[y][x] = sum_s (a_s Wy_s[y]) Wx_s[x],
// i.e. a (32 x S) x (S x 64) matrix product.  From MMA_MIN_HITS spots per region on it runs on the tensor cores
// (mma.sync m16n8k16, bf16 inputs, float32 accumulation): both operands are split into bf16 high and low parts
// and three products (hi hi + hi lo + lo hi) keep ~16 mantissa bits, 2e-5 relative, inside the 1e-4-of-peak
// parity bound.  Sparse regions keep the exact float32 FMA path above.
// ---------------------------------------------------------------------------------------------------
constexpr int MMA_MIN_HITS = 16;

__device__ __forceinline__ float lds32f(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, unsigned v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
// (v0, v1) -> packed bf16 pairs hi and lo with v ~= hi + lo; v0 in the low half (the smaller k index)
__device__ __forceinline__ void bf16_split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void mma_bf16(float &c0, float &c1, float &c2, float &c3, const uint32_t (&a)[4],
                                         uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// Accumulate one warp region; `mma` reports the register layout of the result (see acc_coords).
// hits_s: shared address of this warp's hit list (uint16 [hits_cap]) or 0 when the tensor-core path is off.
//
// Hit-list entry: spot index (12 bits) | x image << 12 | y image << 14.  scipy's mode="reflect" is the sum of
// the direct delta at c and its mirror images at -c - 1 (image 1, spots within R of the low border) and
// 2n - 1 - c (image 2, high border); an image is listed only for regions it reaches, as one more rank-1 term,
// so every tap of the matrix operands is a single table read.
template <bool WIDE, bool DENSE>
__device__ __forceinline__ bool accumulate_region(const RenderParams &p, const FastSmem &s, int n_live, int rx0,
                                                  int ry0, int lane, float (&acc)[8][8], uint32_t hits_s, bool &mma) {
    mma = false;
    if (!DENSE) return accumulate_fast<WIDE>(p, s, n_live, rx0, ry0, lane, acc);  // kernels compiled without the path
    // (the template-level threshold keeps the list building away from patterns too sparse to fill a chunk)
    if (WIDE || hits_s == 0u || n_live < p.mma_tmpl_min) return accumulate_fast<WIDE>(p, s, n_live, rx0, ry0, lane, acc);
    const int R = p.radius;
    const uint32_t spot_s = smem_u32(s.spot);
    const int hits_cap = (p.hits_bytes / RN_WARPS) >> 1;
    // ---- the spots (and mirror images) whose box reaches the region, in list order
    int n_hits = 0;
    for (int base = 0; base < n_live; base += 32) {
        const int j = base + lane;
        unsigned mx = 0, my = 0;  // bit k: image k of this spot reaches the region
        if (j < n_live) {
            const uint2 r = lds64(spot_s + 8u * j);
            const int sx = spot_ix(r), sy = spot_iy(r);
            if (sx + R >= rx0 && sx - R < rx0 + RN_RW && sy + R >= ry0 && sy - R < ry0 + RN_RH) {
                mx = 1u | ((sx < R && rx0 <= R - 1 - sx) ? 2u : 0u) |
                     ((sx >= p.W - R && rx0 + RN_RW - 1 >= 2 * p.W - 1 - sx - R) ? 4u : 0u);
                my = 1u | ((sy < R && ry0 <= R - 1 - sy) ? 2u : 0u) |
                     ((sy >= p.H - R && ry0 + RN_RH - 1 >= 2 * p.H - 1 - sy - R) ? 4u : 0u);
            }
        }
        const int mine = __popc(mx) * __popc(my);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        if (n_hits + total > hits_cap) {  // (cannot happen below ~half the capacity in border spots)
            n_hits = -1;
            break;
        }
        int at = n_hits + incl - mine;
        for (unsigned by = my; by; by &= by - 1)
            for (unsigned bx = mx; bx; bx &= bx - 1)
                sts16(hits_s + 2u * (uint32_t)at++,
                      (unsigned)j | ((unsigned)(__ffs(bx) - 1) << 12) | ((unsigned)(__ffs(by) - 1) << 14));
        n_hits += total;
    }
    __syncwarp();
    if (n_hits < p.mma_min) return accumulate_fast<false>(p, s, n_live, rx0, ry0, lane, acc);
    mma = true;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int g = lane >> 2, t = lane & 3;
    const uint32_t lut0 = smem_u32(s.lut);                  // copy 0: the flat padded kernel L[a]
    const uint32_t lut_hl = lut0 + 64u * (uint32_t)p.n4;    // packed bf16 hi | lo of L[a]
    const int bias = R + LUT_PAD;
    // pixel of B column g in tile 0 and of A row g in tile 0, with the table bias folded in
    const int xb = rx0 + 4 * (g >> 1) + (g & 1) + bias, yb = ry0 + g + bias;
    for (int k0 = 0; k0 < n_hits; k0 += 16) {
        // this lane's four terms of the chunk: k = 2t, 2t + 1, 2t + 8, 2t + 9 (the fragment layout of m16n8k16);
        // the tail is padded with zero-amplitude terms
        uint32_t ax_s[4], ay_s[4];
        float amp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = k0 + 2 * t + (j & 1) + 8 * (j >> 1);
            int cx = rx0, cy = ry0;
            amp[j] = 0.f;
            if (kk < n_hits) {
                const unsigned e = lds16(hits_s + 2u * (uint32_t)kk);
                const uint2 r = lds64(spot_s + 8u * (e & 0xfffu));
                const int sx = spot_ix(r), sy = spot_iy(r);
                const unsigned ix = (e >> 12) & 3u, iy = e >> 14;
                cx = ix == 0u ? sx : (ix == 1u ? -sx - 1 : 2 * p.W - 1 - sx);
                cy = iy == 0u ? sy : (iy == 1u ? -sy - 1 : 2 * p.H - 1 - sy);
                amp[j] = spot_amp(r);
            }
            ax_s[j] = lut_hl + 4u * (uint32_t)(xb - cx);
            ay_s[j] = lut0 + 4u * (uint32_t)(yb - cy);
        }
        // A = a_s Wy_s[y]: rows g, g + 8 of the two 16-row tiles
        uint32_t a_hi[2][4], a_lo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = amp[j] * lds32f(ay_s[j] + 4u * (uint32_t)(16 * mt + 8 * r));
                bf16_split2(v[0], v[1], a_hi[mt][r], a_lo[mt][r]);
                bf16_split2(v[2], v[3], a_hi[mt][2 + r], a_lo[mt][2 + r]);
            }
        // B = Wx_s[x], one 8-column tile at a time; column g of tile nt is pixel 16 (nt >> 1) + 4 (g >> 1) +
        // 2 (nt & 1) + (g & 1), which makes a lane's C columns of a tile pair four consecutive pixels
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = lds32u(ax_s[j] + 4u * (uint32_t)(16 * (nt >> 1) + 2 * (nt & 1)));
            const uint32_t b_hi0 = prmt(w[0], w[1], 0x5410u), b_lo0 = prmt(w[0], w[1], 0x7632u);
            const uint32_t b_hi1 = prmt(w[2], w[3], 0x5410u), b_lo1 = prmt(w[2], w[3], 0x7632u);
            // three products per tile, issued so that consecutive MMAs never share an accumulator
            const int i0 = nt >> 1, i1 = 4 + (nt >> 1), jb = (nt & 1) * 2;
            mma_bf16(acc[i0][jb], acc[i0][jb + 1], acc[i0][jb + 4], acc[i0][jb + 5], a_hi[0], b_hi0, b_hi1);
            mma_bf16(acc[i1][jb], acc[i1][jb + 1], acc[i1][jb + 4], acc[i1][jb + 5], a_hi[1], b_hi0, b_hi1);
            mma_bf16(acc[i0][jb], acc[i0][jb + 1], acc[i0][jb + 4], acc[i0][jb + 5], a_hi[0], b_lo0, b_lo1);
            mma_bf16(acc[i1][jb], acc[i1][jb + 1], acc[i1][jb + 4], acc[i1][jb + 5], a_hi[1], b_lo0, b_lo1);
            mma_bf16(acc[i0][jb], acc[i0][jb + 1], acc[i0][jb + 4], acc[i0][jb + 5], a_lo[0], b_hi0, b_hi1);
            mma_bf16(acc[i1][jb], acc[i1][jb + 1], acc[i1][jb + 4], acc[i1][jb + 5], a_lo[1], b_hi0, b_hi1);
        }
    }
    return true;
}

// Register layouts of a warp region's 64 x 32 pixels (acc[8][8] per lane):
//   LAYOUT_SCALAR  lane (lx = lane & 7, ly = lane >> 3) holds rows ry0 + 8 ly + i, columns
//                  rx0 + 4 lx + (j & 3) + 32 (j >> 2);
//   LAYOUT_MMA     lane (g = lane >> 2, t = lane & 3) holds, in acc[i] with mt = i >> 2, pr = i & 3,
//                  row ry0 + 16 mt + g (j < 4) or that + 8 (j >= 4), columns rx0 + 16 pr + 4 t + (j & 3)
//                  -- the C fragments of mma.m16n8k16 with the column permutation chosen so that a lane
//                  owns four consecutive pixels;
//   LAYOUT_MMA_ST  LAYOUT_MMA after lanes g and g ^ 1 have swapped half of their tiles (mma_to_store_layout):
//                  acc[i] with mt = i >> 2, ph = (i >> 1) & 1, sb = i & 1 is row ry0 + 16 mt + (g & ~1) + sb
//                  (+ 8 for j >= 4), columns rx0 + 16 (2 ph + (g & 1)) + 4 t + (j & 3), so that a store
//                  instruction covers 128 contiguous bytes of 4 rows like the scalar layout does.
enum { LAYOUT_SCALAR = 0, LAYOUT_MMA = 1, LAYOUT_MMA_ST = 2 };

__device__ __forceinline__ void acc_coords(int layout, int lane, int i, int j, int &row, int &col) {
    if (layout == LAYOUT_MMA) {
        row = 16 * (i >> 2) + (lane >> 2) + 8 * (j >> 2);
        col = 16 * (i & 3) + 4 * (lane & 3) + (j & 3);
    } else if (layout == LAYOUT_MMA_ST) {
        const int g = lane >> 2;
        row = 16 * (i >> 2) + (g & ~1) + (i & 1) + 8 * (j >> 2);
        col = 16 * (2 * ((i >> 1) & 1) + (g & 1)) + 4 * (lane & 3) + (j & 3);
    } else {
        row = 8 * (lane >> 3) + i;
        col = 4 * (lane & 7) + (j & 3) + 32 * (j >> 2);
    }
}

// LAYOUT_MMA -> LAYOUT_MMA_ST: even-g lanes hand their odd column tiles to lane ^ 4 and receive its even ones.
__device__ __forceinline__ void mma_to_store_layout(float (&acc)[8][8], int lane) {
    const bool odd = (lane >> 2) & 1;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            const int ia = mt * 4 + 2 * ph, ib = ia + 1;  // column tiles 2 ph and 2 ph + 1 of this lane's rows
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float give = odd ? acc[ia][j] : acc[ib][j];
                const float got = __shfl_xor_sync(0xffffffffu, give, 4);
                // even lane: keeps tile 2 ph of its own row (slot 0), gets tile 2 ph of row g + 1 (slot 1)
                // odd lane:  gets tile 2 ph + 1 of row g - 1 (slot 0), keeps tile 2 ph + 1 of its own row (slot 1)
                if (odd)
                    acc[ia][j] = got;
                else
                    acc[ib][j] = got;
            }
        }
}

// Largest in-frame pixel of a finished region.
template <bool VEC>
__device__ __forceinline__ float region_max(const RenderParams &p, int rx0, int ry0, int lane,
                                            const float (&acc)[8][8], bool mma) {
    const int layout = mma ? LAYOUT_MMA : LAYOUT_SCALAR;
    float m = -INFINITY;
    if (ry0 + RN_RH <= p.H && rx0 + RN_RW <= p.W) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; j += 4)
                m = fmaxf(m, fmaxf(fmaxf(acc[i][j], acc[i][j + 1]), fmaxf(acc[i][j + 2], acc[i][j + 3])));
        return m;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int row, col;
            acc_coords(layout, lane, i, j, row, col);
            if (ry0 + row < p.H && rx0 + col < p.W) m = fmaxf(m, acc[i][j]);
        }
    return m;
}

// Stream one finished warp region to the image (acc * sc, or zeros when no spot reached it) with evict-first
// 16-byte stores: per store instruction a quarter warp covers 128 contiguous bytes of one row (scalar layout;
// 64 bytes in the tensor layout).  Regions entirely inside the frame (all of them when H % 32 == 0 and
// W % 64 == 0) take a branch-free path.
template <bool VEC>
__device__ __forceinline__ void store_region(const RenderParams &p, float *img, int rx0, int ry0, int lane,
                                             float (&acc)[8][8], bool any, float sc, bool mma) {
    if (mma) mma_to_store_layout(acc, lane);
    const int layout = mma ? LAYOUT_MMA_ST : LAYOUT_SCALAR;
    if (VEC && ry0 + RN_RH <= p.H && rx0 + RN_RW <= p.W) {
        const int row4 = p.W >> 2;
        if (mma) {
            int row, col;
            acc_coords(LAYOUT_MMA_ST, lane, 0, 0, row, col);
            float4 *d = reinterpret_cast<float4 *>(img + (size_t)(ry0 + row) * p.W + rx0 + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // acc[i]: 16 rows down per mt, 32 columns right per ph, one row down per sb
                float4 *q = d + (size_t)(16 * (i >> 2) + (i & 1)) * row4 + 8 * ((i >> 1) & 1);
                __stcs(q, make_float4(acc[i][0] * sc, acc[i][1] * sc, acc[i][2] * sc, acc[i][3] * sc));
                __stcs(q + (size_t)8 * row4, make_float4(acc[i][4] * sc, acc[i][5] * sc, acc[i][6] * sc, acc[i][7] * sc));
            }
            return;
        }
        float4 *d = reinterpret_cast<float4 *>(img + (size_t)(ry0 + 8 * (lane >> 3)) * p.W + rx0 + 4 * (lane & 7));
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
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int row, col;
            acc_coords(layout, lane, i, 4 * h, row, col);
            const int y = ry0 + row, x = rx0 + col;
            if (y >= p.H) continue;
            float *d = img + (size_t)y * p.W + x;
            if (VEC) {  // W % 4 == 0: a float4 group is entirely inside or outside the frame
                if (x < p.W)
                    __stcs(reinterpret_cast<float4 *>(d),
                           any ? make_float4(acc[i][4 * h] * sc, acc[i][4 * h + 1] * sc, acc[i][4 * h + 2] * sc,
                                             acc[i][4 * h + 3] * sc)
                               : make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (x + q < p.W) d[q] = any ? acc[i][4 * h + q] * sc : 0.f;
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
