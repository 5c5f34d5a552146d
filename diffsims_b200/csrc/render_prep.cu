// K3a -- per-template preparation for the tcgen05 render kernel (render_umma.cu), one warp per template:
//   float64 projection of the spot rows to detector pixels (simulation2d.py:261-285), in-frame selection and
//   truncation (:422-430), ordering of the live spots by column (ties in list order), "last write wins" inside a
//   pixel (detector_functions.py:297) and the per-half spot lists.
// The work is independent per template, so it runs as a full-GPU pass (thousands of warps) instead of on the one
// front warp of each persistent render CTA, where it was the bottleneck of dense templates (measured: 61 k cycles
// per 680-spot template on a single warp).  The result is one compact record per template in the caller's scratch
// buffer, laid out exactly like a shared-memory slot of the render kernel, which fetches it with one cp.async.bulk:
//   int32 t, n_live, n_half[2], pad[4] | uint2 spot[cap] (column | row << 16, float32 amplitude bits) |
//   uint16 list[2][cap] (indices into spot[] of the spots whose box reaches rows 128 h .. 128 h + 127) |
//   uint32 window[2][ceil(cap / 16)] (first column | columns << 16 that the 16 spots of a list chunk reach)
#include "render_device.cuh"

namespace ds {

constexpr int PREP_BINS = 8 * 33;  // column histogram entries (W <= 256 columns + 1, padded)

__global__ void __launch_bounds__(256) render_prepare_kernel(const RenderParams p, unsigned char *records, const int record_bytes,
                                                             const int warp_bytes, const int window) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    unsigned char *base = smem_raw + (size_t)warp * warp_bytes;
    unsigned *pkey = reinterpret_cast<unsigned *>(base);  // [cap] column | row << 16 of spot j, ~0 = not in frame
    float *pamp = reinterpret_cast<float *>(base + (size_t)p.cap * 4);
    uint2 *sspot = reinterpret_cast<uint2 *>(base + (size_t)p.cap * 8);  // [cap] sorted spots
    int *bins = reinterpret_cast<int *>(base + (size_t)p.cap * 16);
    const int R = p.radius, H = p.H, W = p.W;
    const int n_halves = (H + 127) >> 7;

    for (int t = blockIdx.x * n_warps + warp; t < p.n_tmpl; t += gridDim.x * n_warps) {
        const int n = min(p.count[t], p.cap);
        const double *sxyz = p.xyz + (size_t)t * p.cap * 3;
        const double *sint = p.intensity + (size_t)t * p.cap;
        unsigned char *rec = records + (size_t)t * record_bytes;
        int *hd = reinterpret_cast<int *>(rec);
        uint2 *gspot = reinterpret_cast<uint2 *>(rec + 32);
        unsigned short *glist = reinterpret_cast<unsigned short *>(rec + 32 + (size_t)p.cap * 8);
        auto project = [&](int j) -> unsigned {  // astype(int) truncates
            double px, py;
            project_spot(p, sxyz[3 * j], sxyz[3 * j + 1], px, py);
            if (px >= 0.0 && px < (double)W && py >= 0.0 && py < (double)H) return (unsigned)(int)px | ((unsigned)(int)py << 16);
            return 0xffffffffu;
        };
        // The live spots are ordered by column (ties in list order): the order fixes the float32 sums (reproducible
        // images), keeps the column windows of the render kernel's chunks narrow, and of several spots in one pixel
        // only the last in list order keeps its amplitude.
        int n_live = 0;
        const uint2 *spots;  // where the sorted spots can be read back from
        if (n <= 32) {
            // ---- one spot per lane: rank and overwrite test by all-pairs shuffles
            unsigned kk = 0xffffffffu;
            float a = 0.f;
            if (lane < n) {
                kk = project(lane);
                a = (float)sint[lane];
            }
            const unsigned sk = kk == 0xffffffffu ? 0xffffffffu : (((kk & 0xffffu) << 5) | (unsigned)lane);
            int rank = 0;
            bool dead = false;
            for (int i = 0; i < n; ++i) {
                const unsigned ski = __shfl_sync(0xffffffffu, sk, i), kki = __shfl_sync(0xffffffffu, kk, i);
                rank += ski < sk ? 1 : 0;
                dead |= (kki == kk) & (i > lane);
            }
            n_live = __popc(__ballot_sync(0xffffffffu, kk != 0xffffffffu));
            if (kk != 0xffffffffu) {
                const uint2 v = make_uint2(kk, dead ? 0u : __float_as_uint(a));
                sspot[rank] = v;
                gspot[rank] = v;
            }
            spots = sspot;
        } else {
            // ---- counting sort by column through shared memory
            for (int e = lane; e < PREP_BINS; e += 32) bins[e] = 0;
            __syncwarp();
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                unsigned kk = 0xffffffffu;
                if (j < n) {
                    kk = project(j);
                    if (kk != 0xffffffffu) atomicAdd(&bins[(kk & 0xffffu) + 1], 1);
                    pkey[j] = kk;
                    pamp[j] = (float)sint[j];
                }
                n_live += __popc(__ballot_sync(0xffffffffu, kk != 0xffffffffu));
            }
            __syncwarp();
            {   // exclusive scan of bins[0 .. W]: each lane owns 9 consecutive entries (W + 1 <= 288)
                int loc[9], sum = 0;
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int e = 9 * lane + i;
                    loc[i] = e <= W ? bins[e] : 0;
                    sum += loc[i];
                }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                int run = incl - sum;
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int e = 9 * lane + i;
                    run += loc[i];
                    if (e <= W) bins[e] = run;  // inclusive over bins[0 .. e] = first slot of column e
                }
            }
            __syncwarp();
            // stable scatter: 32 spots at a time, the lanes of one column ranked by lane number
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                const unsigned kk = j < n ? pkey[j] : 0xffffffffu;
                const unsigned mask = __ballot_sync(0xffffffffu, kk != 0xffffffffu);
                int col = 0, at = 0;
                unsigned peers = 0;
                if (kk != 0xffffffffu) {
                    col = (int)(kk & 0xffffu);
                    peers = __match_any_sync(mask, col);
                    at = bins[col] + __popc(peers & ((1u << lane) - 1u));
                }
                __syncwarp();
                if (kk != 0xffffffffu) {
                    sspot[at] = make_uint2(kk, __float_as_uint(pamp[j]));
                    if ((peers >> lane) == 1u) bins[col] = at + 1;  // the highest lane of the column leaves its end
                }
                __syncwarp();
            }
            // last write wins: a spot is overwritten if a later spot of its column (they follow it directly) sits in the
            // same row (lanes only read the other spots' pixel word, which nobody changes)
            for (int i0 = 0; i0 < n_live; i0 += 32) {
                const int i = i0 + lane;
                if (i < n_live) {
                    uint2 v = sspot[i];
                    for (int i2 = i + 1; i2 < n_live; ++i2) {
                        const unsigned k2 = sspot[i2].x;
                        if ((k2 & 0xffffu) != (v.x & 0xffffu)) break;
                        if (k2 == v.x) {
                            v.y = 0u;
                            break;
                        }
                    }
                    sspot[i].y = v.y;
                    gspot[i] = v;
                }
            }
            spots = sspot;
        }
        __syncwarp();
        // ---- per-half lists: the spots whose box reaches rows [128 h, 128 h + 127] (the folded images of a spot lie
        // inside its own clipped box); overwritten spots are left out
        int n_half[2] = {0, 0};
        for (int h = 0; h < n_halves; ++h) {
            unsigned short *list = glist + (size_t)h * p.cap;
            int cnt = 0;
            for (int j0 = 0; j0 < n_live; j0 += 32) {
                const int j = j0 + lane;
                bool hit = false;
                if (j < n_live) {
                    const uint2 r = spots[j];
                    const int sy = (int)(r.x >> 16);
                    hit = r.y != 0u && sy + R >= 128 * h && sy - R <= 128 * h + 127;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (hit) list[cnt + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)j;
                cnt += __popc(mask);
            }
            n_half[h] = cnt;
            // column window of every chunk of 16 list entries, 16-column aligned (the first chunk of a half covers, and
            // thereby zeroes, every column of the accumulator)
            __syncwarp();
            unsigned *gwin = reinterpret_cast<unsigned *>(rec + umma_windows_offset(p.cap)) + (size_t)h * ((p.cap + 15) / 16);
            const int Wp = (W + 15) & ~15;
            for (int c = lane; 16 * c < max(cnt, 1); c += 32) {
                int col0 = 0, ncols = Wp;
                if (c > 0 && window) {
                    int xmin = 1 << 20, xmax = -1;
                    for (int e = 16 * c; e < min(16 * c + 16, cnt); ++e) {
                        const int sx = (int)(spots[list[e]].x & 0xffffu);
                        xmin = min(xmin, sx);
                        xmax = max(xmax, sx);
                    }
                    col0 = max(0, xmin - R) & ~15;
                    ncols = ((min(W, xmax + R + 1) - col0) + 15) & ~15;
                }
                gwin[c] = (unsigned)col0 | ((unsigned)ncols << 16);
            }
        }
        if (lane == 0) {
            hd[0] = t;
            hd[1] = n_live;
            hd[2] = n_half[0];
            hd[3] = n_half[1];
        }
        __syncwarp();
    }
}

int launch_render_prepare(const RenderParams &p, unsigned char *records, int window, cudaStream_t st) {
    const int record_bytes = umma_record_bytes(p.cap);
    const int warp_bytes = (p.cap * 16 + PREP_BINS * 4 + 15) & ~15;
    int warps = 8;
    while (warps > 1 && (size_t)warps * warp_bytes > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * warp_bytes;
    if (smem > 48 * 1024) cudaFuncSetAttribute(render_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int want = (p.n_tmpl + warps - 1) / warps;
    const int grid = want < 16 * num_sms() ? want : 16 * num_sms();
    render_prepare_kernel<<<grid, warps * 32, smem, st>>>(p, records, record_bytes, warp_bytes, window);
    return check_launch("ds_render (prepare)");
}

}  // namespace ds
