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

}  // namespace ds

extern "C" int ds_structure_factors(void *stream, int32_t n_g, const double *hkl, const double *gnorm,
                                    int32_t n_atoms, const double *frac, const double *occ, int32_t n_elem,
                                    const int32_t *elem_start, const double *coeffs, const double *dw,
                                    int32_t scattering_model, const double *prefactor, double *F_out,
                                    double *I_out) {
    using namespace ds;
    DS_REQUIRE(n_g >= 0 && n_atoms >= 0 && n_elem >= 0, "ds_structure_factors: negative size");
    DS_REQUIRE(n_elem <= SF_MAX_ELEM, "ds_structure_factors: more than %d distinct elements", SF_MAX_ELEM);
    DS_REQUIRE(scattering_model >= 0 && scattering_model <= 2, "ds_structure_factors: unknown scattering model %d",
               scattering_model);
    if (n_g == 0) return 0;
    const dim3 grid((n_g + SF_THREADS - 1) / SF_THREADS), block(SF_THREADS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
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
