"""``DiffractionLibraryGenerator`` -- template library over a ``StructureLibrary``
(diffsims/generators/library_generator.py:36-152), B200-native: one K2 launch per phase instead of a
Python loop over orientations.  ``VectorLibraryGenerator`` is a different algorithm and out of scope."""
import numpy as np
import torch

from .. import engine
from ..library import pack_csr
from ..libraries.diffraction_library import DiffractionLibrary, LazyObjectArray
from ..sims.diffraction_simulation import DiffractionSimulation

__all__ = ["DiffractionLibraryGenerator"]


class DiffractionLibraryGenerator:
    def __init__(self, electron_diffraction_calculator):
        self.electron_diffraction_calculator = electron_diffraction_calculator

    def get_diffraction_library(self, structure_library, calibration, reciprocal_radius, half_shape,
                                with_direct_beam=True, max_excitation_error=1e-2, shape_factor_width=None,
                                debye_waller_factors={}):
        """Dictionary of diffraction data for every structure and orientation of the library; same
        parameters, keys and attribute names as the reference (:51-152)."""
        diffraction_library = DiffractionLibrary()
        diffractor = self.electron_diffraction_calculator
        if shape_factor_width is None:
            shape_factor_width = max_excitation_error
        for phase_name in structure_library.struct_lib.keys():
            structure, orientations = structure_library.struct_lib[phase_name]
            num_orientations = len(orientations)
            gt, spots = diffractor.calculate_ed_data_batch(
                structure, reciprocal_radius, orientations, max_excitation_error, shape_factor_width,
                debye_waller_factors)
            # padded rows -> CSR on the device (ds_pack_csr), pixel coordinates of the packed rows (:129-132,
            # ds_library_pixel_coords over the packed list as one long row), one compact device->host copy; the
            # per-orientation objects are made lazily (a 3e5-orientation library costs no Python loop here)
            packed = pack_csr(spots)
            total = packed.g_index.shape[0]
            pix = engine.library_pixel_coords(torch.tensor([total], dtype=torch.int32, device=packed.xyz.device),
                                              packed.xyz.reshape(1, max(total, 1), 3) if total else packed.xyz.reshape(1, 0, 3),
                                              calibration, half_shape).reshape(-1, 2).cpu().numpy() if total else \
                np.zeros((0, 2), dtype=np.int32)
            off = packed.offsets.cpu().numpy()
            xyz = packed.xyz.cpu().numpy()
            inten = packed.intensity.cpu().numpy()
            hkl = gt.hkl[packed.g_index.cpu().numpy()]

            def make_simulation(i, off=off, xyz=xyz, inten=inten, hkl=hkl):
                lo, hi = off[i], off[i + 1]
                sim = DiffractionSimulation(coordinates=xyz[lo:hi].copy(), indices=hkl[lo:hi], intensities=inten[lo:hi].copy(),
                                            with_direct_beam=with_direct_beam)
                sim.calibration = calibration
                return sim

            simulations = LazyObjectArray(num_orientations, make_simulation)
            # the direct-beam mask of the container (with_direct_beam=False hides the (000) row, sims/...:171-179) applies to
            # the pixel and intensity lists as well
            mask_of = (lambda lo, hi, xyz=xyz: np.ones(hi - lo, dtype=bool)) if with_direct_beam else \
                (lambda lo, hi, xyz=xyz: np.any(xyz[lo:hi], axis=1))
            pixel_coords = LazyObjectArray(num_orientations, lambda i, off=off, pix=pix, m=mask_of:
                                           pix[off[i]:off[i + 1]][m(off[i], off[i + 1])].astype(int))
            intensities = LazyObjectArray(num_orientations, lambda i, off=off, inten=inten, m=mask_of:
                                          inten[off[i]:off[i + 1]][m(off[i], off[i + 1])].copy())

            diffraction_library[phase_name] = {
                "simulations": simulations,
                "orientations": orientations,
                "pixel_coords": pixel_coords,
                "intensities": intensities,
            }
        diffraction_library.identifiers = structure_library.identifiers
        diffraction_library.structures = structure_library.structures
        diffraction_library.diffraction_generator = diffractor
        diffraction_library.reciprocal_radius = reciprocal_radius
        diffraction_library.with_direct_beam = with_direct_beam
        return diffraction_library
