"""``DiffractionGenerator`` -- the OLD api simulator (diffsims/generators/diffraction_generator.py:108-331),
B200-native.  ``calculate_ed_data`` keeps its signature and returns a ``DiffractionSimulation``; the
batched ``calculate_ed_data_batch`` (used by ``DiffractionLibraryGenerator``) simulates every
orientation of a structure with one launch of the K2 kernel.

Differences from the new api that are preserved: plain structure in its own
(diffpy) orientation, Euler angles rzxz in degrees, g set = index box +-floor(rr / |a*|, |b*|, |c*|)
filtered ``|g| < rr`` (strict) INCLUDING a single (000), integer ``indices``.
Profile simulations and ``AtomicDiffractionGenerator`` are out of scope.
"""
import numpy as np

from .. import engine
from ..crystal import Rotation
from ..sims.diffraction_simulation import DiffractionSimulation
from ..utils import shape_factor_models as sfm
from ..utils.sim_utils import get_electron_wavelength, get_points_in_sphere

__all__ = ["DiffractionGenerator"]

_shape_factor_model_mapping = {
    "linear": sfm.linear,
    "atanc": sfm.atanc,
    "sinc": sfm.sinc,
    "sin2c": sfm.sin2c,
    "lorentzian": sfm.lorentzian,
}


class DiffractionGenerator(object):
    """Computes electron diffraction patterns for a crystal structure (kinematical)."""

    def __init__(self, accelerating_voltage, scattering_params="lobato", precession_angle=0,
                 shape_factor_model="lorentzian", approximate_precession=True, minimum_intensity=1e-20,
                 **kwargs):
        self.wavelength = get_electron_wavelength(accelerating_voltage)
        self.precession_angle = np.abs(precession_angle)
        self.approximate_precession = approximate_precession
        if isinstance(shape_factor_model, str):
            if shape_factor_model in _shape_factor_model_mapping.keys():
                self.shape_factor_model = _shape_factor_model_mapping[shape_factor_model]
            else:
                raise NotImplementedError(
                    f"{shape_factor_model} is not a recognized shape factor "
                    f"model, choose from: {_shape_factor_model_mapping.keys()} "
                    f"or provide your own function.")
        else:
            self.shape_factor_model = shape_factor_model
        self.minimum_intensity = minimum_intensity
        self.shape_factor_kwargs = kwargs
        if scattering_params in ["lobato", "xtables", None]:
            self.scattering_params = scattering_params
        else:
            raise NotImplementedError(
                "The scattering parameters `{}` is not implemented. "
                "See documentation for available "
                "implementations.".format(scattering_params))

    def _native_model(self):
        minima = float(self.shape_factor_kwargs.get("minima_number", 5))
        if self.precession_angle != 0 and self.approximate_precession:
            return "lorentzian_precession", minima
        name = sfm.NATIVE.get(self.shape_factor_model)
        if name is None or not set(self.shape_factor_kwargs) <= {"minima_number"}:
            raise NotImplementedError(
                "the old api runs native shape factor models only "
                "(binary, linear, sinc, sin2c, atanc, lorentzian)")
        return name, minima

    def _g_table(self, structure, reciprocal_radius, debye_waller_factors):
        recip = structure.lattice.reciprocal()
        idx, cart, _ = get_points_in_sphere(recip, reciprocal_radius)  # recomputed per call in the reference
        return engine.make_gtable(structure, idx.astype(np.int64), cart, debye_waller_factors,
                                  self.scattering_params)

    def calculate_ed_data_batch(self, structure, reciprocal_radius, rotations, max_excitation_error=1e-2,
                                shape_factor_width=None, debye_waller_factors={}):
        """All orientations (sequence of Euler rzxz triples in degrees) at once: returns
        (engine.GTable, engine.SpotTable) with the reflections in the reference's order."""
        gt = self._g_table(structure, reciprocal_radius, debye_waller_factors)
        if shape_factor_width is None:
            shape_factor_width = max_excitation_error
        eul = np.asarray(rotations, dtype=float).reshape(-1, 3)
        # R = euler2mat(rzxz) = Rz(phi1) Rx(Phi) Rz(phi2) (diffraction_generator.py:247-253) is the active
        # matrix of the inverse Bunge quaternion
        quats = (~Rotation.from_euler(eul, degrees=True)).data
        model, minima = self._native_model()
        spots = engine.simulate(gt, quats, self.wavelength, max_excitation_error, shape_factor_width, model,
                                minima, float(np.deg2rad(self.precession_angle)), self.minimum_intensity)
        return gt, spots

    def calculate_ed_data(self, structure, reciprocal_radius, rotation=(0, 0, 0), with_direct_beam=True,
                          max_excitation_error=1e-2, shape_factor_width=None, debye_waller_factors={}):
        """Electron diffraction data of ``structure`` for one orientation (:192-331)."""
        gt, spots = self.calculate_ed_data_batch(structure, reciprocal_radius, [rotation],
                                                 max_excitation_error, shape_factor_width,
                                                 debye_waller_factors)
        n = int(spots.count[0])
        idx = spots.g_index[0, :n].cpu().numpy()
        return DiffractionSimulation(coordinates=spots.xyz[0, :n].cpu().numpy(), indices=gt.hkl[idx],
                                     intensities=spots.intensity[0, :n].cpu().numpy(),
                                     with_direct_beam=with_direct_beam)
