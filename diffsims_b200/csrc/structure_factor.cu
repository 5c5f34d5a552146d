// K1 -- kinematical structure factors F(g), evaluated once per (phase, g-set).
//
// Replaces, for the whole reciprocal-lattice table at once, what the reference recomputes for every
// rotation: _get_kinematical_structure_factor (diffsims/utils/sim_utils.py:256-304),
// get_atomic_scattering_factors (:227-253) and the prefactor*|F|^2 of get_kinematical_intensities (:353).
//
// One thread per reciprocal-lattice vector; the atom table is staged through shared memory in tiles
// (x, y, z, occupancy as doubles).  Arithmetic is float64 throughout: the Lobato parameterisation has
// cancelling terms (|a_i| ~ 200 summing to ~1) and the parity contract is rtol 1e-5 on |F|^2 including
// weak reflections, which float32 phases/sums cannot meet (hard part 2).  The kernel is
// therefore bound by the FP64 pipe (sincospi ~ 40 DFMA per atom x g pair), not by HBM or the SFU.
//
// Large cells with integer Miller indices (the per-phase g table) take the FACTORISED kernels instead:
// exp(2 pi i (h x + k y + l z)) = Ex_j[h] Ey_j[k] Ez_j[l], so a small pre-pass tabulates the three phase factors of every
// atom for h, k, l in [-H, H] (3 N_at (2 H + 1) sincospi in all, the occupancy folded into Ex) and the main kernel needs
// two complex multiplications per atom x g pair -- 8 DFMA instead of ~40 -- reading table tiles staged in shared memory
// (lanes of a warp hold consecutive l: Ex / Ey reads are broadcasts, Ez reads are contiguous).
#include <stdarg.h>

#include "common.cuh"

namespace ds {

constexpr int SF_THREADS = 128;
constexpr int SF_ATOM_TILE = 1024;
constexpr int SF_MAX_ELEM = 128;

template <int MODEL>
__device__ __forceinline__ double scattering_factor(double g2, const double *__restrict__ c /*[5][2]*/) {
    if (MODEL == DS_SCATT_NONE) return 1.0;
    double f = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double a = c[2 * i], t = g2 * c[2 * i + 1];
        if (MODEL == DS_SCATT_LOBATO) {
            const double d = 1.0 + t;
            f += a * ((2.0 + t) * (1.0 / (d * d)));  // sim_utils.py:248
        } else {
            f += a * exp(-0.25 * t);  // sim_utils.py:250
        }
    }
    return f;
}

template <int MODEL>
__global__ void __launch_bounds__(SF_THREADS)
structure_factor_kernel(int n_g, const double *__restrict__ hkl, const double *__restrict__ gnorm, int n_atoms,
                        const double *__restrict__ frac, const double *__restrict__ occ, int n_elem,
                        const int *__restrict__ elem_start, const double *__restrict__ coeffs,
                        const double *__restrict__ dw, const double *__restrict__ prefactor,
                        double *__restrict__ F_out, double *__restrict__ I_out) {
    __shared__ double4 s_atom[SF_ATOM_TILE];  // x, y, z, occupancy
    __shared__ double s_coef[SF_MAX_ELEM * 10];
    __shared__ double s_dw[SF_MAX_ELEM];
    __shared__ int s_start[SF_MAX_ELEM + 1];

    for (int i = threadIdx.x; i < n_elem * 10; i += SF_THREADS) s_coef[i] = coeffs[i];
    for (int i = threadIdx.x; i < n_elem; i += SF_THREADS) s_dw[i] = dw[i];
    for (int i = threadIdx.x; i <= n_elem; i += SF_THREADS) s_start[i] = elem_start[i];

    const int g = blockIdx.x * SF_THREADS + threadIdx.x;
    const bool live = g < n_g;
    double h = 0, k = 0, l = 0, g2 = 0;
    if (live) {
        h = hkl[3 * g + 0];
        k = hkl[3 * g + 1];
        l = hkl[3 * g + 2];
        const double gn = gnorm[g];
        g2 = gn * gn;
    }
    double Fre = 0.0, Fim = 0.0;

    for (int base = 0; base < n_atoms; base += SF_ATOM_TILE) {
        __syncthreads();
        const int n_tile = min(SF_ATOM_TILE, n_atoms - base);
        for (int i = threadIdx.x; i < n_tile; i += SF_THREADS) {
            const int j = base + i;
            s_atom[i] = make_double4(frac[3 * j], frac[3 * j + 1], frac[3 * j + 2], occ[j]);
        }
        __syncthreads();
        for (int e = 0; e < n_elem; ++e) {
            const int lo = max(s_start[e], base), hi = min(s_start[e + 1], base + n_tile);
            if (lo >= hi) continue;  // uniform across the CTA
            // f_e(g^2) * exp(-g^2 B_e / 4): the real part of the reference's complex exponent (:297-301)
            const double fe = scattering_factor<MODEL>(g2, &s_coef[e * 10]) * exp(-0.25 * g2 * s_dw[e]);
            double re = 0.0, im = 0.0;
            for (int j = lo - base; j < hi - base; ++j) {
                const double4 a = s_atom[j];
                const double ph = h * a.x + k * a.y + l * a.z;  // hkl . r_j in turns
                double sn, cs;
                sincospi(2.0 * ph, &sn, &cs);
                re = fma(a.w, cs, re);
                im = fma(a.w, sn, im);
            }
            Fre = fma(fe, re, Fre);
            Fim = fma(fe, im, Fim);
        }
    }
    if (live) {
        if (F_out) {
            F_out[2 * g] = Fre;
            F_out[2 * g + 1] = Fim;
        }
        if (I_out) {
            const double p = prefactor ? prefactor[g] : 1.0;
            I_out[g] = p * (Fre * Fre + Fim * Fim);  // sim_utils.py:353
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// factorised phases (integer hkl, |h|, |k|, |l| <= H)
// ---------------------------------------------------------------------------------------------------
constexpr int SFT_THREADS = 256;
constexpr int SFT_G_PER_THREAD = 1;
constexpr int SFT_G_TILE = SFT_THREADS * SFT_G_PER_THREAD;

// table[axis][atom][m + H] = (occ_j if axis == 0 else 1) * exp(2 pi i m r_j[axis]),  m = -H .. H
__global__ void sf_phase_table_kernel(int n_atoms, int H, const double *__restrict__ frac, const double *__restrict__ occ,
                                      double2 *__restrict__ table) {
    const int W = 2 * H + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3ll * n_atoms * W) return;
    const int m = (int)(idx % W) - H, j = (int)((idx / W) % n_atoms), axis = (int)(idx / ((long long)W * n_atoms));
    double sn, cs;
    sincospi(2.0 * (double)m * frac[3 * j + axis], &sn, &cs);
    const double w = axis == 0 ? occ[j] : 1.0;
    table[idx] = make_double2(w * cs, w * sn);
}

template <int MODEL>
__global__ void __launch_bounds__(SFT_THREADS)
structure_factor_tab_kernel(int n_g, const double *__restrict__ hkl, const double *__restrict__ gnorm, int n_atoms, int n_elem,
                            const int *__restrict__ elem_start, const double *__restrict__ coeffs, const double *__restrict__ dw,
                            const double *__restrict__ prefactor, int H, int smem_entries, const double2 *__restrict__ table,
                            double *__restrict__ F_out, double *__restrict__ I_out) {
    extern __shared__ __align__(16) unsigned char sft_smem[];
    double2 *s_tab = reinterpret_cast<double2 *>(sft_smem);  // [atom_tile][Wh + Wk + Wl]: Ex | Ey | Ez of an atom
    __shared__ double s_coef[SF_MAX_ELEM * 10];
    __shared__ double s_dw[SF_MAX_ELEM];
    __shared__ int s_start[SF_MAX_ELEM + 1];
    __shared__ int s_lo[3], s_hi[3];
    const int W = 2 * H + 1;
    for (int i = threadIdx.x; i < n_elem * 10; i += SFT_THREADS) s_coef[i] = coeffs[i];
    for (int i = threadIdx.x; i < n_elem; i += SFT_THREADS) s_dw[i] = dw[i];
    for (int i = threadIdx.x; i <= n_elem; i += SFT_THREADS) s_start[i] = elem_start[i];
    if (threadIdx.x < 3) {
        s_lo[threadIdx.x] = 1 << 20;
        s_hi[threadIdx.x] = -(1 << 20);
    }
    __syncthreads();

    // this CTA's g vectors (consecutive table rows: l runs fastest, so h and k span a few values only) and the index
    // ranges its table tiles need
    int ih[SFT_G_PER_THREAD], ik[SFT_G_PER_THREAD], il[SFT_G_PER_THREAD];
    double g2[SFT_G_PER_THREAD], Fre[SFT_G_PER_THREAD], Fim[SFT_G_PER_THREAD];
    bool live[SFT_G_PER_THREAD];
#pragma unroll
    for (int q = 0; q < SFT_G_PER_THREAD; ++q) {
        const int g = blockIdx.x * SFT_G_TILE + q * SFT_THREADS + threadIdx.x;
        live[q] = g < n_g;
        ih[q] = ik[q] = il[q] = 0;
        g2[q] = 0.0;
        if (live[q]) {
            ih[q] = (int)hkl[3 * g + 0];
            ik[q] = (int)hkl[3 * g + 1];
            il[q] = (int)hkl[3 * g + 2];
            g2[q] = gnorm[g] * gnorm[g];
            atomicMin(&s_lo[0], ih[q]), atomicMax(&s_hi[0], ih[q]);
            atomicMin(&s_lo[1], ik[q]), atomicMax(&s_hi[1], ik[q]);
            atomicMin(&s_lo[2], il[q]), atomicMax(&s_hi[2], il[q]);
        }
        Fre[q] = Fim[q] = 0.0;
    }
    __syncthreads();
    const int h0 = s_lo[0], k0 = s_lo[1], l0 = s_lo[2];
    const int Wh = s_hi[0] - h0 + 1, Wk = s_hi[1] - k0 + 1, Wl = s_hi[2] - l0 + 1;
    const int rows = Wh + Wk + Wl;
    const int atom_tile = max(1, min(128, smem_entries / rows));  // (rows <= 3 (2 H + 1) <= smem_entries: at least one atom)
#pragma unroll
    for (int q = 0; q < SFT_G_PER_THREAD; ++q) {  // offsets inside an atom's smem row (idle threads read entry 0)
        ih[q] = live[q] ? ih[q] - h0 : 0;
        ik[q] = live[q] ? Wh + ik[q] - k0 : 0;
        il[q] = live[q] ? Wh + Wk + il[q] - l0 : 0;
    }
    for (int e = 0; e < n_elem; ++e) {
        double fe[SFT_G_PER_THREAD], re[SFT_G_PER_THREAD], im[SFT_G_PER_THREAD];
#pragma unroll
        for (int q = 0; q < SFT_G_PER_THREAD; ++q) {
            // f_e(g^2) * exp(-g^2 B_e / 4): the real part of the reference's complex exponent (:297-301)
            fe[q] = scattering_factor<MODEL>(g2[q], &s_coef[e * 10]) * exp(-0.25 * g2[q] * s_dw[e]);
            re[q] = im[q] = 0.0;
        }
        for (int base = s_start[e]; base < s_start[e + 1]; base += atom_tile) {
            const int n_tile = min(atom_tile, s_start[e + 1] - base);
            __syncthreads();  // the previous tile has been consumed
            for (int i = threadIdx.x; i < n_tile * rows; i += SFT_THREADS) {
                const int j = i / rows, r = i % rows;
                // row r of an atom: Ex[h0 + r] | Ey[k0 + r - Wh] | Ez[l0 + r - Wh - Wk]
                const int axis = r < Wh ? 0 : (r < Wh + Wk ? 1 : 2);
                const int m = axis == 0 ? h0 + r : (axis == 1 ? k0 + r - Wh : l0 + r - Wh - Wk);
                s_tab[i] = table[((long long)axis * n_atoms + base + j) * W + m + H];
            }
            __syncthreads();
            for (int j = 0; j < n_tile; ++j) {
                const double2 *row = s_tab + (size_t)j * rows;
#pragma unroll
                for (int q = 0; q < SFT_G_PER_THREAD; ++q) {
                    const double2 a = row[ih[q]], b = row[ik[q]], c = row[il[q]];
                    const double pr = a.x * b.x - a.y * b.y, pi = a.x * b.y + a.y * b.x;
                    re[q] += pr * c.x - pi * c.y;
                    im[q] += pr * c.y + pi * c.x;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SFT_G_PER_THREAD; ++q) {
            Fre[q] = fma(fe[q], re[q], Fre[q]);
            Fim[q] = fma(fe[q], im[q], Fim[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < SFT_G_PER_THREAD; ++q) {
        const int g = blockIdx.x * SFT_G_TILE + q * SFT_THREADS + threadIdx.x;
        if (!live[q]) continue;
        if (F_out) {
            F_out[2 * g] = Fre[q];
            F_out[2 * g + 1] = Fim[q];
        }
        if (I_out) {
            const double p = prefactor ? prefactor[g] : 1.0;
            I_out[g] = p * (Fre[q] * Fre[q] + Fim[q] * Fim[q]);  // sim_utils.py:353
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// factorised phases over the index BOX (dense integer tables of large cells)
//
// For a fixed h and element e,  S_e[k][l] = sum_{j in e} (Ex_j[h] Ey_j[k]) Ez_j[l]  is a small complex matrix product over
// the atoms.  A CTA takes one h, 16 values of k and 64 of l; a thread keeps a 4 (k) x 2 (l) register tile, so an atom costs
// it six 16-byte shared loads for 32 DFMA (one complex multiply-add per pair: 4 DFMA instead of the 8 of the row kernel
// above, and no per-pair index arithmetic) -- bound by the FP64 pipe.  Box entries that are not rows of the table cost
// nothing when a whole warp tile (8 x 32) is empty, and are dropped at the end otherwise.  The atoms are split over up to
// eight CTAs per tile (a 61^3 box has only ~130 live CTA tiles); every split writes its partial sums to its own box-shaped
// scratch array and a gather pass adds them in a fixed order into the table rows (duplicates of an index triple, like
// the second (000) the new API appends, read the same entries).
// ---------------------------------------------------------------------------------------------------
constexpr int SFB_MAX_H = 63;
// result boxes (= atom units) of the box kernel: enough CTAs to fill the GPU for cells of a few hundred atoms, <= 64 MB
inline int sfb_splits(int n_atoms, int H) {
    const long long W = 2ll * H + 1, box_bytes = W * W * W * 16;
    long long s = (n_atoms + 31) / 32;
    if (s > 16) s = 16;
    if (s * box_bytes > (64ll << 20)) s = (64ll << 20) / box_bytes;
    return s < 1 ? 1 : (int)s;
}
constexpr int SFB_THREADS = 128;   // 4 warps: 2 (k) x 2 (l)
constexpr int SFB_KT = 16, SFB_LT = 64;
constexpr int SFB_ATOMS = 32;      // atoms per shared-memory tile: 32 x (16 + 64) x 16 B = 40 KB

__global__ void sf_box_clear_kernel(long long n, int *__restrict__ idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = -1;
}
__global__ void sf_box_scatter_kernel(int n_g, const double *__restrict__ hkl, int H, int *__restrict__ idx) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_g) return;
    const int W = 2 * H + 1;
    const int h = (int)hkl[3 * g] + H, k = (int)hkl[3 * g + 1] + H, l = (int)hkl[3 * g + 2] + H;
    idx[((long long)h * W + k) * W + l] = g;  // (rows sharing an index triple: any of them, they share |g|)
}

// Units of the atom split: every unit is a run of at most `unit` atoms of ONE element (so that the element's scattering
// factor can be applied after the sum), `unit` the smallest multiple of the tile size for which the units fit the result
// boxes.  Returns the number of units; with u >= 0 also unit u's element and atom range.  (n_elem <= max_units.)
__device__ __forceinline__ int sfb_units(const int *start, int n_elem, int max_units, int u, int &e_out, int &lo, int &hi) {
    int unit = SFB_ATOMS, n_units;
    for (;; unit += SFB_ATOMS) {
        n_units = 0;
        for (int e = 0; e < n_elem; ++e) n_units += (start[e + 1] - start[e] + unit - 1) / unit;
        if (n_units <= max_units) break;
    }
    e_out = lo = hi = 0;
    int at = 0;
    for (int e = 0; e < n_elem && u >= 0; ++e) {
        const int n_e = start[e + 1] - start[e], pieces = (n_e + unit - 1) / unit;
        if (u < at + pieces) {
            e_out = e;
            lo = start[e] + (u - at) * unit;
            hi = min(start[e + 1], lo + unit);
            break;
        }
        at += pieces;
    }
    return n_units;
}

__global__ void __launch_bounds__(SFB_THREADS)
structure_factor_box_kernel(int n_atoms, int n_elem, const int *__restrict__ elem_start, int H, int max_units, int n_ltiles,
                            const double2 *__restrict__ table, const int *__restrict__ idx, double2 *__restrict__ Sbox) {
    extern __shared__ __align__(16) unsigned char sfb_smem[];
    double2 *s_P = reinterpret_cast<double2 *>(sfb_smem);        // [SFB_ATOMS][SFB_KT]: Ex_j[h] Ey_j[k]
    double2 *s_Z = s_P + SFB_ATOMS * SFB_KT;                     // [SFB_ATOMS][SFB_LT]: Ez_j[l]
    __shared__ int s_start[SF_MAX_ELEM + 1];
    const int W = 2 * H + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wk = warp >> 1, wl = warp & 1, lk = lane >> 4, ll = lane & 15;
    const int h = blockIdx.y, k_cta = blockIdx.x * SFB_KT, l_cta = (blockIdx.z % n_ltiles) * SFB_LT;
    const int unit = blockIdx.z / n_ltiles;
    // this thread's entries: k = k_cta + kq + rk (rk = 0..3), l = l_cta + lq + 16 rl (rl = 0, 1)
    const int kq = wk * 8 + lk * 4, lq = wl * 32 + ll;
    bool live[4][2], any = false;
#pragma unroll
    for (int rk = 0; rk < 4; ++rk)
#pragma unroll
        for (int rl = 0; rl < 2; ++rl) {
            const int k = k_cta + kq + rk, l = l_cta + lq + 16 * rl;
            live[rk][rl] = (k < W && l < W) && __ldg(idx + ((long long)h * W + k) * W + l) >= 0;
            any |= live[rk][rl];
        }
    const bool warp_live = __any_sync(0xffffffffu, any);
    if (!__syncthreads_or(warp_live)) return;  // no table row in this CTA's part of the box
    for (int i = tid; i <= n_elem; i += SFB_THREADS) s_start[i] = elem_start[i];
    __syncthreads();
    int e_unit, atom_lo, atom_hi;
    if (unit >= sfb_units(s_start, n_elem, max_units, unit, e_unit, atom_lo, atom_hi)) return;  // (uniform) unused unit
    const double2 *tab_x = table, *tab_y = table + (size_t)n_atoms * W, *tab_z = table + 2 * (size_t)n_atoms * W;
    double re[4][2], im[4][2];
#pragma unroll
    for (int rk = 0; rk < 4; ++rk)
#pragma unroll
        for (int rl = 0; rl < 2; ++rl) re[rk][rl] = im[rk][rl] = 0.0;
    for (int base = atom_lo; base < atom_hi; base += SFB_ATOMS) {
        const int n_tile = min(SFB_ATOMS, atom_hi - base);
        __syncthreads();  // the previous tile has been consumed
        for (int i = tid; i < n_tile * SFB_KT; i += SFB_THREADS) {
            const int j = i / SFB_KT, k = k_cta + i % SFB_KT;
            double2 v = make_double2(0.0, 0.0);
            if (k < W) {
                const double2 ex = __ldg(tab_x + (size_t)(base + j) * W + h), ey = __ldg(tab_y + (size_t)(base + j) * W + k);
                v = make_double2(ex.x * ey.x - ex.y * ey.y, ex.x * ey.y + ex.y * ey.x);
            }
            s_P[i] = v;
        }
        for (int i = tid; i < n_tile * SFB_LT; i += SFB_THREADS) {
            const int j = i / SFB_LT, l = l_cta + i % SFB_LT;
            s_Z[i] = l < W ? __ldg(tab_z + (size_t)(base + j) * W + l) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (!warp_live) continue;  // (warp-uniform; the barriers above are still taken)
        for (int j = 0; j < n_tile; ++j) {
            const double2 *pr = s_P + j * SFB_KT + kq, *zr = s_Z + j * SFB_LT + lq;
            const double2 z0 = zr[0], z1 = zr[16];
#pragma unroll
            for (int rk = 0; rk < 4; ++rk) {
                const double2 a = pr[rk];
                re[rk][0] = fma(a.x, z0.x, re[rk][0]);
                re[rk][0] = fma(-a.y, z0.y, re[rk][0]);
                im[rk][0] = fma(a.x, z0.y, im[rk][0]);
                im[rk][0] = fma(a.y, z0.x, im[rk][0]);
                re[rk][1] = fma(a.x, z1.x, re[rk][1]);
                re[rk][1] = fma(-a.y, z1.y, re[rk][1]);
                im[rk][1] = fma(a.x, z1.y, im[rk][1]);
                im[rk][1] = fma(a.y, z1.x, im[rk][1]);
            }
        }
    }
    double2 *out = Sbox + (size_t)unit * W * W * W;
#pragma unroll
    for (int rk = 0; rk < 4; ++rk)
#pragma unroll
        for (int rl = 0; rl < 2; ++rl) {
            if (!live[rk][rl]) continue;
            const int k = k_cta + kq + rk, l = l_cta + lq + 16 * rl;
            out[((long long)h * W + k) * W + l] = make_double2(re[rk][rl], im[rk][rl]);
        }
}

// F(row) = sum over the units, in order, of f_e(g^2) exp(-g^2 B_e / 4) S_unit[row's box entry]; the scattering factor of an
// element is evaluated once per row
template <int MODEL>
__global__ void __launch_bounds__(128)
sf_box_gather_kernel(int n_g, const double *__restrict__ hkl, const double *__restrict__ gnorm, int H, int n_elem,
                     const int *__restrict__ elem_start, const double *__restrict__ coeffs, const double *__restrict__ dw,
                     int max_units, const double2 *__restrict__ Sbox, const double *__restrict__ prefactor,
                     double *__restrict__ F_out, double *__restrict__ I_out) {
    __shared__ double s_coef[SF_MAX_ELEM * 10];
    __shared__ double s_dw[SF_MAX_ELEM];
    __shared__ int s_start[SF_MAX_ELEM + 1];
    __shared__ int s_unit_elem[32];
    __shared__ int s_n_units;
    for (int i = threadIdx.x; i < n_elem * 10; i += blockDim.x) s_coef[i] = coeffs[i];
    for (int i = threadIdx.x; i < n_elem; i += blockDim.x) s_dw[i] = dw[i];
    for (int i = threadIdx.x; i <= n_elem; i += blockDim.x) s_start[i] = elem_start[i];
    __syncthreads();
    if (threadIdx.x < 32) {
        int e, lo, hi;
        const int n_units = sfb_units(s_start, n_elem, max_units, threadIdx.x, e, lo, hi);
        s_unit_elem[threadIdx.x] = e;
        if (threadIdx.x == 0) s_n_units = n_units;
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_g) return;
    const int W = 2 * H + 1;
    const int h = (int)hkl[3 * g] + H, k = (int)hkl[3 * g + 1] + H, l = (int)hkl[3 * g + 2] + H;
    const size_t entry = ((size_t)h * W + k) * W + l, box = (size_t)W * W * W;
    const double g2 = gnorm[g] * gnorm[g];
    double Fre = 0.0, Fim = 0.0, fe = 0.0;
    int e_cur = -1;
    for (int u = 0; u < s_n_units; ++u) {  // fixed order: reproducible sums (units are ordered by element)
        const int e = s_unit_elem[u];
        if (e != e_cur) {
            // f_e(g^2) * exp(-g^2 B_e / 4): the real part of the reference's complex exponent (sim_utils.py:297-301)
            fe = scattering_factor<MODEL>(g2, &s_coef[e * 10]) * exp(-0.25 * g2 * s_dw[e]);
            e_cur = e;
        }
        const double2 part = Sbox[(size_t)u * box + entry];
        Fre = fma(fe, part.x, Fre);
        Fim = fma(fe, part.y, Fim);
    }
    if (F_out) {
        F_out[2 * g] = Fre;
        F_out[2 * g + 1] = Fim;
    }
    if (I_out) {
        const double p = prefactor ? prefactor[g] : 1.0;
        I_out[g] = p * (Fre * Fre + Fim * Fim);  // sim_utils.py:353
    }
}

}  // namespace ds

extern "C" int64_t ds_structure_factors_scratch_bytes(int32_t n_atoms, int32_t hkl_int_max) {
    if (n_atoms < 0 || hkl_int_max < 0) return -1;
    const long long W = 2ll * hkl_int_max + 1;
    // phase-factor tables | (box kernel, H <= 63) index box (int32) | result box (complex128)
    long long bytes = 3ll * n_atoms * W * 16;
    if (hkl_int_max <= ds::SFB_MAX_H) bytes += ((W * W * W * 4 + 15) & ~15ll) + ds::sfb_splits(n_atoms, hkl_int_max) * W * W * W * 16;
    return bytes;
}

extern "C" int ds_structure_factors(void *stream, int32_t n_g, const double *hkl, const double *gnorm,
                                    int32_t n_atoms, const double *frac, const double *occ, int32_t n_elem,
                                    const int32_t *elem_start, const double *coeffs, const double *dw,
                                    int32_t scattering_model, const double *prefactor, double *F_out,
                                    double *I_out, int32_t hkl_int_max, void *table_scratch) {
    using namespace ds;
    DS_REQUIRE(n_g >= 0 && n_atoms >= 0 && n_elem >= 0, "ds_structure_factors: negative size");
    DS_REQUIRE(n_elem <= SF_MAX_ELEM, "ds_structure_factors: more than %d distinct elements", SF_MAX_ELEM);
    DS_REQUIRE(scattering_model >= 0 && scattering_model <= 2, "ds_structure_factors: unknown scattering model %d",
               scattering_model);
    if (n_g == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // factorised path: integer indices in a bounded range, a cell large enough to pay for the table pre-pass
    if (hkl_int_max > 0 && hkl_int_max <= 127 && table_scratch != nullptr && n_atoms >= 32 && n_g >= 4096 &&
        (reinterpret_cast<uintptr_t>(table_scratch) & 15) == 0) {
        const int H = hkl_int_max, W = 2 * H + 1;
        const size_t smem = 32 * 1024;
        const int smem_entries = (int)(smem / 16);  // >= 3 W for H <= 127
        double2 *table = static_cast<double2 *>(table_scratch);
        const long long n_tab = 3ll * n_atoms * W;
        sf_phase_table_kernel<<<(unsigned)((n_tab + 255) / 256), 256, 0, st>>>(n_atoms, H, frac, occ, table);
        // tables that fill a good part of their index box: the register-tiled box kernel
        const long long box = (long long)W * W * W;
        const int max_units = sfb_splits(n_atoms, H);
        if (H <= SFB_MAX_H && box <= 16ll * n_g && n_elem <= max_units) {
            unsigned char *after = static_cast<unsigned char *>(table_scratch) + n_tab * 16;
            int *idx = reinterpret_cast<int *>(after);
            double2 *Sbox = reinterpret_cast<double2 *>(after + ((box * 4 + 15) & ~15ll));
            sf_box_clear_kernel<<<(unsigned)((box + 255) / 256), 256, 0, st>>>(box, idx);
            sf_box_scatter_kernel<<<(n_g + 255) / 256, 256, 0, st>>>(n_g, hkl, H, idx);
            const int n_ltiles = (W + SFB_LT - 1) / SFB_LT;
            const dim3 grid_b((W + SFB_KT - 1) / SFB_KT, W, n_ltiles * max_units);
            const size_t smem_b = (size_t)SFB_ATOMS * (SFB_KT + SFB_LT) * 16;
            cudaFuncSetAttribute(structure_factor_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            structure_factor_box_kernel<<<grid_b, SFB_THREADS, smem_b, st>>>(n_atoms, n_elem, elem_start, H, max_units, n_ltiles, table,
                                                                             idx, Sbox);
#define DS_SFG_LAUNCH(M)                                                                                                      \
    sf_box_gather_kernel<M><<<(n_g + 127) / 128, 128, 0, st>>>(n_g, hkl, gnorm, H, n_elem, elem_start, coeffs, dw, max_units, \
                                                                Sbox, prefactor, F_out, I_out)
            if (scattering_model == DS_SCATT_LOBATO)
                DS_SFG_LAUNCH(DS_SCATT_LOBATO);
            else if (scattering_model == DS_SCATT_XTABLES)
                DS_SFG_LAUNCH(DS_SCATT_XTABLES);
            else
                DS_SFG_LAUNCH(DS_SCATT_NONE);
#undef DS_SFG_LAUNCH
            return check_launch("ds_structure_factors (box)");
        }
        const int grid_t = (n_g + SFT_G_TILE - 1) / SFT_G_TILE;
#define DS_SFT_LAUNCH(M)                                                                                                   \
    do {                                                                                                                   \
        if (smem > 48 * 1024)                                                                                              \
            cudaFuncSetAttribute(structure_factor_tab_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
        structure_factor_tab_kernel<M><<<grid_t, SFT_THREADS, smem, st>>>(n_g, hkl, gnorm, n_atoms, n_elem, elem_start,    \
                                                                          coeffs, dw, prefactor, H, smem_entries, table,   \
                                                                          F_out, I_out);                                   \
    } while (0)
        if (scattering_model == DS_SCATT_LOBATO)
            DS_SFT_LAUNCH(DS_SCATT_LOBATO);
        else if (scattering_model == DS_SCATT_XTABLES)
            DS_SFT_LAUNCH(DS_SCATT_XTABLES);
        else
            DS_SFT_LAUNCH(DS_SCATT_NONE);
#undef DS_SFT_LAUNCH
        return check_launch("ds_structure_factors (factorised)");
    }
    const dim3 grid((n_g + SF_THREADS - 1) / SF_THREADS), block(SF_THREADS);
#define DS_SF_LAUNCH(M)                                                                                       \
    structure_factor_kernel<M><<<grid, block, 0, st>>>(n_g, hkl, gnorm, n_atoms, frac, occ, n_elem, elem_start, \
                                                       coeffs, dw, prefactor, F_out, I_out)
    if (scattering_model == DS_SCATT_LOBATO)
        DS_SF_LAUNCH(DS_SCATT_LOBATO);
    else if (scattering_model == DS_SCATT_XTABLES)
        DS_SF_LAUNCH(DS_SCATT_XTABLES);
    else
        DS_SF_LAUNCH(DS_SCATT_NONE);
#undef DS_SF_LAUNCH
    return check_launch("ds_structure_factors");
}
