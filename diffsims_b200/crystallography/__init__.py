"""Reciprocal-lattice containers of the hot path (host-side bookkeeping only).

Mirrors the parts of diffsims/crystallography/reciprocal_lattice_vector.py and
diffsims/crystallography/_diffracting_vector.py that the template-simulation path
touches; symmetry bookkeeping, structure-factor methods,
printing etc. are out of scope.
"""
from __future__ import annotations

import copy
import itertools

import numpy as np

from ..crystal import Rotation

__all__ = ["ReciprocalLatticeVector", "DiffractingVector", "g_set_from_min_dspacing"]


def _highest_hkl(lattice, min_dspacing):
    # orix.vector.miller._get_highest_hkl, called at reciprocal_lattice_vector.py:1131
    highest = np.ones(3, dtype=int)
    for axis in range(3):
        hkl = np.zeros(3)
        d = min_dspacing + 1
        while d > min_dspacing:
            hkl[axis] += 1
            d = 1 / lattice.rnorm(hkl)
        highest[axis] = hkl[axis]
    return highest


def g_set_from_min_dspacing(lattice, min_dspacing, include_zero_vector=False):
    """Integer (hkl) with d >= min_dspacing in the reference's order, (000) appended last.

    reciprocal_lattice_vector.py:1077-1142: index box from the highest (h00), (0k0), (00l) reached,
    descending lexicographic order, (000) removed, inclusive cut d >= min_dspacing.
    """
    hi = _highest_hkl(lattice, min_dspacing)
    box = np.asarray(list(itertools.product(*[np.arange(-i, i + 1) for i in hi])), dtype=np.int64)
    box = box[~np.all(box == 0, axis=1)][::-1]
    keep = 1 / lattice.rnorm(box) >= min_dspacing
    hkl = box[keep]
    if include_zero_vector:
        hkl = np.vstack([hkl, np.zeros((1, 3), dtype=np.int64)])
    return np.ascontiguousarray(hkl)


class ReciprocalLatticeVector:
    """Reciprocal lattice vectors (hkl) of a phase, stored as Cartesian ``data`` [n, 3]."""

    def __init__(self, phase, xyz=None, hkl=None, hkil=None):
        self._phase_factory = None
        if callable(phase) and not hasattr(phase, "structure"):
            # lazily built phase (the rotated-basis copy of a packed simulation result)
            self._phase, self._phase_factory = None, phase
            if xyz is None:
                raise ValueError("a lazy phase needs Cartesian `xyz`")
        else:
            self._phase = phase
        if self._phase is not None and getattr(phase, "point_group", None) is None:
            raise ValueError(f"The phase {phase} must have a point group set")
        if sum(v is not None for v in (xyz, hkl, hkil)) != 1:
            raise ValueError("Exactly one of `xyz`, `hkl`, or `hkil` must be passed")
        self._coordinate_format = "hkl"
        self._hkl_exact = None
        if xyz is not None:
            data = np.atleast_2d(np.asarray(xyz))
        else:
            if hkil is not None:
                hkil = np.atleast_2d(np.asarray(hkil, dtype=float))
                if not np.allclose(hkil[:, :3].sum(axis=1), 0, atol=1e-4):
                    raise ValueError("The Miller-Bravais indices convention h + k + i = 0 is not satisfied")
                hkl = hkil[:, [0, 1, 3]]
                self._coordinate_format = "hkil"
            hkl = np.atleast_2d(np.asarray(hkl))
            data = hkl.astype(float) @ np.asarray(phase.structure.lattice.recbase).T
        # vectors are kept flat; N-D inputs are flattened in the order of orix `Object3d.flatten` (first axis fastest),
        # which is what the reference's `to_flat_polar` sees (_diffracting_vector.py:186-194)
        self.data = data.T.reshape(3, -1).T if data.ndim > 2 else data.reshape(-1, 3)

    @property
    def phase(self):
        if self._phase is None:
            self._phase = self._phase_factory()
        return self._phase

    @phase.setter
    def phase(self, value):
        self._phase = value

    # -- sizes --------------------------------------------------------------
    @property
    def size(self):
        return self.data.shape[0]

    @property
    def shape(self):
        return (self.data.shape[0],)

    def __len__(self):
        return self.size

    # -- coordinates ----------------------------------------------------------
    @property
    def hkl(self):
        """Miller indices: Cartesian -> reciprocal, ``data @ base.T`` (:174)."""
        if self._hkl_exact is not None:
            return self._hkl_exact.astype(float)
        return self.data @ np.asarray(self.phase.structure.lattice.base).T

    @property
    def h(self):
        return self.hkl[..., 0]

    @property
    def k(self):
        return self.hkl[..., 1]

    @property
    def l(self):
        return self.hkl[..., 2]

    @property
    def hkil(self):
        hkl = self.hkl
        return np.stack([hkl[:, 0], hkl[:, 1], -(hkl[:, 0] + hkl[:, 1]), hkl[:, 2]], axis=1)

    @property
    def coordinate_format(self):
        return self._coordinate_format

    @property
    def coordinates(self):
        return getattr(self, self._coordinate_format)

    @property
    def gspacing(self):
        """|g| in 1/Angstrom (:440 ``lattice.rnorm(hkl)``)."""
        return np.sqrt((np.asarray(self.data, dtype=float) ** 2).sum(axis=-1))

    @property
    def dspacing(self):
        return 1 / self.gspacing

    @property
    def has_hexagonal_lattice(self):
        a, b, c, al, be, ga = self.phase.structure.lattice.abcABG()
        return bool(np.isclose(a, b) and np.allclose([al, be, ga], [90, 90, 120]))

    @classmethod
    def from_min_dspacing(cls, phase, min_dspacing=0.7, include_zero_vector=False):
        hkl = g_set_from_min_dspacing(phase.structure.lattice, min_dspacing, include_zero_vector)
        new = cls(phase, hkl=hkl)
        new._hkl_exact = hkl
        return new

    def __getitem__(self, key):
        ph = self._phase if self._phase is not None else self._phase_factory
        new = self.__class__(ph, xyz=np.atleast_2d(self.data[key]))
        if self._hkl_exact is not None:
            new._hkl_exact = np.atleast_2d(self._hkl_exact[key])
        return new

    def flatten(self):
        return self

    def deepcopy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        name = self.__class__.__name__
        data = np.array_str(self.coordinates, precision=0, suppress_small=True)
        return f"{name} {self.shape}, {self.phase.name} ({self.phase.point_group})\n{data}"


class DiffractingVector(ReciprocalLatticeVector):
    """Reflections that intersect the Ewald sphere, with intensities
    (diffsims/crystallography/_diffracting_vector.py:25-194)."""

    def __init__(self, phase, xyz=None, hkl=None, hkil=None, intensity=None):
        super().__init__(phase, xyz=xyz, hkl=hkl, hkil=hkil)
        if intensity is None:
            self._intensity = np.full(self.shape, np.nan)
        elif len(intensity) != self.size:
            raise ValueError("Length of intensity array must match number of vectors")
        else:
            self._intensity = np.array(intensity)

    def __getitem__(self, key):
        new = super().__getitem__(key)
        if np.isnan(np.asarray(self.intensity, dtype=float)).all():
            new._intensity = np.full(new.shape, np.nan)
        else:
            sl = self.intensity[key]
            if not hasattr(sl, "__len__"):
                sl = np.array([sl])
            new._intensity = sl
        return new

    def __setitem__(self, key, value):
        # orix Object3d.__setitem__ writes into ``data`` (used at simulation2d.py:283-284)
        self.data[key] = value

    @property
    def basis_rotation(self):
        return Rotation.from_matrix(self.phase.structure.lattice.baserot)

    def rotate_with_basis(self, rotation):
        """Rotate vectors and lattice basis (:127-161): data @ G with G = rotation.to_matrix()."""
        if rotation.size != 1:
            raise ValueError("Rotation must be a single rotation")
        G = np.asarray(rotation.to_matrix()).reshape(3, 3)
        new_phase = self.phase.deepcopy()
        lat = new_phase.structure.lattice
        lat.setLatPar(baserot=np.asarray(lat.baserot) @ G)
        new = ReciprocalLatticeVector(new_phase, xyz=np.asarray(self.data, dtype=float) @ G)
        new._hkl_exact = self._hkl_exact
        return new

    @property
    def intensity(self):
        return self._intensity

    @intensity.setter
    def intensity(self, value):
        if not hasattr(value, "__len__"):
            value = np.array([value] * self.size)
        if len(value) != self.size:
            raise ValueError("Length of intensity array must match number of vectors")
        self._intensity = np.array(value)

    def calculate_structure_factor(self):
        raise NotImplementedError(
            "Structure factor calculation not implemented for DiffractionVector. "
            "Use ReciprocalLatticeVector instead.")

    def to_flat_polar(self):
        """(r, theta) of the vectors projected on the x-y plane (:186-194)."""
        d = np.asarray(self.data, dtype=float)
        return np.linalg.norm(d[:, :2], axis=1), np.arctan2(d[:, 1], d[:, 0])
