"""``DiffractionLibraryGenerator`` -- template library over a ``StructureLibrary``
(diffsims/generators/library_generator.py:36-152), B200-native: one K2 launch per phase instead of a
Python loop over orientations.  ``VectorLibraryGenerator`` is a different algorithm and out of scope."""
import numpy as np

from .. import engine
from ..libraries.diffraction_library import DiffractionLibrary
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
            pix = engine.library_pixel_coords(spots.count, spots.xyz, calibration, half_shape).cpu().numpy()
            count = spots.count.cpu().numpy()
            xyz = spots.xyz.cpu().numpy()
            inten = spots.intensity.cpu().numpy()
            gidx = spots.g_index.cpu().numpy()

            simulations = np.empty(num_orientations, dtype="object")
            pixel_coords = np.empty(num_orientations, dtype="object")
            intensities = np.empty(num_orientations, dtype="object")
            for i in range(num_orientations):
                n = count[i]
                simulation = DiffractionSimulation(
                    coordinates=xyz[i, :n].copy(), indices=gt.hkl[gidx[i, :n]], intensities=inten[i, :n].copy(),
                    with_direct_beam=with_direct_beam)
                simulation.calibration = calibration
                simulations[i] = simulation
                # :129-132, computed for the whole library by ds_library_pixel_coords; the direct-beam mask of
                # the container (with_direct_beam=False hides the (000) row) applies to the pixel list as well
                pixel_coords[i] = pix[i, :n][simulation.direct_beam_mask].astype(int)
                intensities[i] = simulation.intensities

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
