"""``DiffractionLibrary`` (diffsims/libraries/diffraction_library.py:29-178): dict of per-phase
simulations / orientations / pixel coordinates / intensities, with the reference's pickle io."""
import pickle

import numpy as np

__all__ = ["DiffractionLibrary", "load_DiffractionLibrary"]


def load_DiffractionLibrary(filename, safety=False):
    if safety:
        with open(filename, "rb") as handle:
            return pickle.load(handle)
    raise RuntimeError("Unpickling is risky, turn safety to True if you trust the author of this content")


def _get_library_entry_from_angles(library, phase, angles):
    """First entry whose Euler angles are within 1e-2 (summed absolute difference) of ``angles``."""
    for orientation_index, orientation in enumerate(library[phase]["orientations"]):
        if np.sum(np.abs(np.subtract(orientation, angles))) < 1e-2:
            return orientation_index
    raise ValueError("It appears that no library entry lies with 1e-2 of the target angle")


class DiffractionLibrary(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.identifiers = None
        self.structures = None
        self.diffraction_generator = None
        self.reciprocal_radius = 0.0
        self.with_direct_beam = False

    def get_library_entry(self, phase=None, angle=None):
        if phase is not None:
            phase_entry = self[phase]
            orientation_index = _get_library_entry_from_angles(self, phase, angle) if angle is not None else 0
        elif angle is not None:
            raise ValueError("To select a certain angle you must first specify a phase")
        else:
            phase_entry = next(iter(self.values()))
            orientation_index = 0
        return {
            "Sim": phase_entry["simulations"][orientation_index],
            "intensities": phase_entry["intensities"][orientation_index],
            "pixel_coords": phase_entry["pixel_coords"][orientation_index],
            "pattern_norm": np.linalg.norm(phase_entry["intensities"][orientation_index]),
        }

    def pickle_library(self, filename):
        with open(filename, "wb") as handle:
            pickle.dump(self, handle, protocol=pickle.HIGHEST_PROTOCOL)
