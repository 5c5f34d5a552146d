"""Rel-rod shape-factor models: names, native ids and user-facing callables.

The arithmetic used by simulations lives in the simulate kernel (csrc/simulate.cu,
``shape_factor``); these Python callables exist so that user code written against
diffsims/utils/shape_factor_models.py keeps working (they can be passed as
``shape_factor_model=`` and are mapped back to the native model) and so that a result
returned with excitation errors can be post-processed on the host.
"""
import numpy as np

__all__ = ["atanc", "binary", "linear", "lorentzian", "lorentzian_precession", "sin2c", "sinc"]


def binary(excitation_error, max_excitation_error):
    """Unit weight for every reflection that intersects (reference :33-49)."""
    return 1


def linear(excitation_error, max_excitation_error):
    """max(0, 1 - |s| / s_max) (reference :52-73)."""
    sf = 1 - np.abs(excitation_error) / max_excitation_error
    return np.maximum(sf, 0.0) if isinstance(sf, np.ndarray) else max(sf, 0.0)


def sinc(excitation_error, max_excitation_error, minima_number=5):
    """|sin(x) / x|, x = pi n s / s_max; 0 at s == 0 as in the reference (:76-101)."""
    x = np.asarray(np.pi * minima_number / max_excitation_error * excitation_error, dtype=float)
    out = np.zeros_like(x)
    nz = x != 0
    out[nz] = np.abs(np.sin(x[nz]) / x[nz])
    return out


def sin2c(excitation_error, max_excitation_error, minima_number=5):
    """sinc squared (reference :104-123)."""
    return sinc(excitation_error, max_excitation_error, minima_number) ** 2


def atanc(excitation_error, max_excitation_error, minima_number=5):
    """atan(x) / x, x = pi n s / |s_max|; 1 at s == 0 (reference :126-151)."""
    x = np.asarray(np.pi * minima_number / np.abs(max_excitation_error) * excitation_error, dtype=float)
    out = np.ones_like(x)
    nz = x != 0
    out[nz] = np.arctan(x[nz]) / x[nz]
    return out


def lorentzian(excitation_error, max_excitation_error):
    """Two-beam rocking-curve approximation, Palatinus et al. (2019) eq. 6 (reference :154-180)."""
    sigma = np.pi / max_excitation_error
    return sigma / (np.pi * (sigma ** 2 * excitation_error ** 2 + 1)) * max_excitation_error


def lorentzian_precession(excitation_error, max_excitation_error, r_spot, precession_angle):
    """Precessed Lorentzian, Palatinus et al. (2019) eq. 10 (reference :183-219)."""
    sigma = np.pi / max_excitation_error
    u = sigma ** 2 * (r_spot ** 2 * precession_angle ** 2 - excitation_error ** 2) + 1
    z = np.sqrt(u ** 2 + 4 * sigma ** 2 * excitation_error ** 2)
    return (sigma / np.pi) * np.sqrt(2 * (u + z) / z ** 2)


# native model name for each callable above (the kernel evaluates these)
NATIVE = {binary: "binary", linear: "linear", sinc: "sinc", sin2c: "sin2c", atanc: "atanc",
          lorentzian: "lorentzian"}
