"""Leaf functions of diffsims/utils/sim_utils.py that sit on the template-simulation path."""
import itertools
import math

import numpy as np

from .. import engine
from ..engine import get_scattering_params_dict  # noqa: F401  (re-export, sim_utils.py:139)

__all__ = ["get_electron_wavelength", "get_kinematical_intensities", "get_points_in_sphere",
           "get_scattering_params_dict", "is_lattice_hexagonal"]

# CODATA constants as scipy.constants exposes them (the reference imports h, m_e, e, c)
try:
    from scipy.constants import c as _c, e as _e, h as _h, m_e as _m_e
except Exception:  # pragma: no cover
    _h, _m_e, _e, _c = 6.62607015e-34, 9.1093837139e-31, 1.602176634e-19, 299792458.0


def get_electron_wavelength(accelerating_voltage):
    """Relativistic electron wavelength in Angstrom for a voltage in kV (sim_utils.py:57-79)."""
    if accelerating_voltage in (np.inf, "inf"):
        return 0
    E = accelerating_voltage * 1e3
    return _h / math.sqrt(2 * _m_e * _e * E * (1 + (_e / (2 * _m_e * _c * _c)) * E)) * 1e10


def get_kinematical_intensities(structure, g_indices, g_hkls_array, debye_waller_factors=None,
                                scattering_params="lobato", prefactor=1):
    """Peak intensities prefactor * |F(g)|^2 (sim_utils.py:307-354), evaluated by the structure-factor
    kernel (K1) on the GPU.  Returns a float64 numpy array like the reference."""
    g_indices = np.asarray(g_indices)
    if g_indices.size == 0:
        return np.zeros(0)
    _, I = engine.structure_factors(structure, g_indices, g_hkls_array, debye_waller_factors,
                                    scattering_params, prefactor=None if np.isscalar(prefactor) else prefactor,
                                    want_F=False)
    out = I.cpu().numpy()
    if np.isscalar(prefactor) and prefactor != 1:
        out = prefactor * out
    return out


def get_kinematical_structure_factor(structure, g_indices, g_hkls_array, debye_waller_factors=None,
                                     scattering_params="lobato"):
    """Complex F(g) (sim_utils.py:256-304) from K1, as a complex128 numpy array."""
    F, _ = engine.structure_factors(structure, g_indices, g_hkls_array, debye_waller_factors,
                                    scattering_params, want_I=False)
    F = F.cpu().numpy()
    return F[:, 0] + 1j * F[:, 1]


def get_points_in_sphere(reciprocal_lattice, reciprocal_radius):
    """All reciprocal lattice points with |g| < reciprocal_radius (strict) inside the index box
    +-floor(radius / |a*|, |b*|, |c*|) -- host-side enumeration of the OLD api (sim_utils.py:436-474).

    Returns (indices [n,3] int, cartesian [n,3], distances [n])."""
    a, b, c = reciprocal_lattice.a, reciprocal_lattice.b, reciprocal_lattice.c
    rng = [np.arange(-np.floor(reciprocal_radius / v), np.floor(reciprocal_radius / v) + 1) for v in (a, b, c)]
    pts = np.asarray(list(itertools.product(*rng)))
    dist = reciprocal_lattice.dist(pts, [0, 0, 0])
    keep = np.abs(dist) < reciprocal_radius
    idx = pts[keep]
    return idx, reciprocal_lattice.cartesian(idx), dist[keep]


def is_lattice_hexagonal(latt):
    """True for hexagonal / trigonal-hex lattices (sim_utils.py:477-494)."""
    truth = latt.a == latt.b
    truth = truth and latt.alpha == 90
    truth = truth and latt.beta == 90
    truth = truth and latt.gamma == 120
    return truth
