"""``DiffractionLibrary`` -- per-phase template store of the OLD api.

Behavioural mirror of diffsims/libraries/diffraction_library.py:29-178: a ``dict`` keyed by phase name whose
values hold the object arrays ``simulations`` / ``orientations`` / ``pixel_coords`` / ``intensities``, plus the
library-level attributes the generator fills in, ``get_library_entry`` and pickle persistence (with the
reference's explicit ``safety`` opt-in for loading).
"""
import pickle

import numpy as np

__all__ = ["DiffractionLibrary", "load_DiffractionLibrary", "LazyObjectArray"]

_ANGLE_TOLERANCE = 1e-2  # summed |delta Euler| below which an orientation counts as found (reference :62)


def load_DiffractionLibrary(filename, safety=False):
    """Unpickle a library saved with ``pickle_library``; refuses unless ``safety=True``."""
    if not safety:
        raise RuntimeError("Unpickling is risky, turn safety to True if you trust the author of this content")
    with open(filename, "rb") as fh:
        return pickle.load(fh)


def _get_library_entry_from_angles(library, phase, angles):
    """Index of the first orientation of ``phase`` within the angle tolerance of ``angles`` (ValueError if none)."""
    target = np.asarray(angles, dtype=float)
    for index, euler in enumerate(library[phase]["orientations"]):
        if np.abs(np.asarray(euler, dtype=float) - target).sum() < _ANGLE_TOLERANCE:
            return index
    raise ValueError("It appears that no library entry lies with 1e-2 of the target angle")


def _as_object_array(items):
    out = np.empty(len(items), dtype="object")
    for i, it in enumerate(items):
        out[i] = it
    return out


class LazyObjectArray:
    """Stand-in for the reference's 1-D numpy object arrays (``simulations`` / ``pixel_coords`` / ``intensities``) over
    the packed result of a batched build: entry ``i`` is made by ``make(i)`` when it is first asked for, so a library
    of 3e5 orientations costs no per-orientation Python work until it is looked at.  Indexing with an integer returns
    the entry, with a slice / index array a real object array; ``np.asarray`` and pickling materialise everything."""

    dtype = np.dtype("object")
    ndim = 1

    def __init__(self, n, make):
        self._n, self._make, self._cache = int(n), make, {}

    def __len__(self):
        return self._n

    @property
    def shape(self):
        return (self._n,)

    @property
    def size(self):
        return self._n

    def _one(self, i):
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(f"index {i} is out of bounds for axis 0 with size {self._n}")
        if i not in self._cache:
            self._cache[i] = self._make(i)
        return self._cache[i]

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self._one(key)
        if isinstance(key, slice):
            return _as_object_array([self._one(i) for i in range(*key.indices(self._n))])
        idx = np.asarray(key)
        if idx.dtype == bool:
            idx = np.nonzero(idx)[0]
        return _as_object_array([self._one(i) for i in idx.reshape(-1)])

    def __iter__(self):
        return (self._one(i) for i in range(self._n))

    def __array__(self, dtype=None, copy=None):
        return _as_object_array(list(self))

    def __reduce__(self):           # pickles as the plain object array the reference stores
        return (_as_object_array, (list(self),))


class DiffractionLibrary(dict):
    """Maps phase name -> simulated diffraction data for every orientation of that phase."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.identifiers = self.structures = self.diffraction_generator = None
        self.reciprocal_radius = 0.0
        self.with_direct_beam = False

    def get_library_entry(self, phase=None, angle=None):
        """One entry as a dict with keys ``Sim``, ``intensities``, ``pixel_coords``, ``pattern_norm``.

        Without ``phase`` the first phase is used (an ``angle`` then makes no sense -> ValueError); without
        ``angle`` the first orientation."""
        if phase is None:
            if angle is not None:
                raise ValueError("To select a certain angle you must first specify a phase")
            entry, index = next(iter(self.values())), 0
        else:
            entry = self[phase]
            index = 0 if angle is None else _get_library_entry_from_angles(self, phase, angle)
        intensities = entry["intensities"][index]
        return {"Sim": entry["simulations"][index], "intensities": intensities,
                "pixel_coords": entry["pixel_coords"][index], "pattern_norm": np.linalg.norm(intensities)}

    def pickle_library(self, filename):
        """Save with the highest pickle protocol (load with ``load_DiffractionLibrary(..., safety=True)``)."""
        with open(filename, "wb") as fh:
            pickle.dump(self, fh, protocol=pickle.HIGHEST_PROTOCOL)
