"""``DiffractionSimulation`` -- result container of the OLD api
(diffsims/sims/diffraction_simulation.py:32-354): coordinates / indices / intensities of one pattern,
calibration, the direct-beam mask and ``get_diffraction_pattern`` (rendered by the K3 kernel).
Masks, plotting and ``ProfileSimulation`` are out of scope."""
import copy

import numpy as np
import torch

from .. import engine

__all__ = ["DiffractionSimulation"]


class DiffractionSimulation:
    def __init__(self, coordinates, indices=None, intensities=None, calibration=None, offset=(0.0, 0.0),
                 with_direct_beam=False):
        coordinates = np.asarray(coordinates)
        if coordinates.ndim == 1:
            coordinates = coordinates[None, :]
        if indices is None:
            indices = np.full((coordinates.shape[0], 3), np.nan)
        if intensities is None:
            intensities = np.full((coordinates.shape[0]), np.nan)
        if (coordinates.shape[0] == indices.shape[0] == intensities.shape[0]
                and coordinates.ndim == indices.ndim == 2 and intensities.ndim == 1):
            self._coordinates = coordinates
            self._indices = indices
            self._intensities = intensities
        else:
            raise ValueError(
                "Coordinate, intensity, and indices lists must be of the correct and matching shape.")
        self.calibration = calibration
        self.offset = np.array(offset)
        self.with_direct_beam = with_direct_beam

    def __len__(self):
        return self.coordinates.shape[0]

    @property
    def size(self):
        return self.__len__()

    def __getitem__(self, sliced):
        coords = self.coordinates[sliced]
        inds = self.indices[sliced]
        ints = self.intensities[sliced]
        if coords.ndim == 1:
            coords, inds, ints = coords[None, :], inds[None, :], ints[None]
        if coords.ndim > 2 or coords.shape[1] > 3 or coords.shape[1] < 2:
            raise ValueError(f"Invalid slice: {sliced}")
        return DiffractionSimulation(coords, indices=inds, intensities=ints, calibration=self.calibration,
                                     offset=self.offset, with_direct_beam=self.with_direct_beam)

    def deepcopy(self):
        return copy.deepcopy(self)

    def __add__(self, other):
        new = self.deepcopy()
        new.extend(other)
        return new

    def extend(self, other):
        self._coordinates = np.concatenate([self._coordinates, other._coordinates], axis=0)
        self._indices = np.concatenate([self._indices, other._indices], axis=0)
        self._intensities = np.concatenate([self._intensities, other._intensities], axis=0)

    @property
    def direct_beam_mask(self):
        """Rows exposed by ``coordinates`` / ``indices`` / ``intensities``: everything when
        ``with_direct_beam``, else every row except the all-zero (000) one (reference :171-179)."""
        if self.with_direct_beam:
            return np.ones(self._intensities.shape, dtype=bool)
        return self._coordinates.any(axis=1)

    def _masked_view(name):  # noqa: N805  (class-body helper: masked read / masked write of one raw array)
        raw = "_" + name

        def getter(self):
            return getattr(self, raw)[self.direct_beam_mask]

        def setter(self, value):
            getattr(self, raw)[self.direct_beam_mask] = value

        return property(getter, setter, doc=f"The {name} of all unmasked points.")

    indices = _masked_view("indices")
    coordinates = _masked_view("coordinates")
    intensities = _masked_view("intensities")
    del _masked_view

    @property
    def calibrated_coordinates(self):
        """Coordinates in pixels (:143-149)."""
        if self.calibration is not None:
            return (self.coordinates[:, :2] + self.offset) / self.calibration
        raise Exception("Pixel calibration is not set!")

    @property
    def calibration(self):
        return self._calibration

    @calibration.setter
    def calibration(self, value):
        if value is not None:
            if np.all(np.equal(value, 0)):
                raise ValueError("`calibration` cannot be zero.")
            if isinstance(value, (float, int)):
                value = (value, value)
            elif len(value) != 2:
                raise ValueError("`calibration` must be a float or length-2tuple of floats.")
            value = np.array(value)
        self._calibration = value

    def _get_transformed_coordinates(self, angle, center=(0, 0), mirrored=False, units="real"):
        """Translate, rotate or mirror the spot coordinates (:199-215)."""
        c = self.coordinates.copy() if units == "real" else self.calibrated_coordinates.copy()
        cx, cy = center
        x, y = c[:, 0].copy(), c[:, 1].copy()
        theta = (-1 if mirrored else 1) * np.arctan2(y, x) + np.deg2rad(angle)
        rd = np.sqrt(x ** 2 + y ** 2)
        c[:, 0] = rd * np.cos(theta) + cx
        c[:, 1] = rd * np.sin(theta) + cy
        return c

    def rotate_shift_coordinates(self, angle, center=(0, 0), mirrored=False):
        self.coordinates = self._get_transformed_coordinates(angle, center, mirrored, units="real")

    def get_diffraction_pattern(self, shape=(512, 512), sigma=10, direct_beam_position=None,
                                in_plane_angle=0, mirrored=False):
        """Normalised pattern with a Gaussian per reflection (:296-354), rasterised by K3.

        The reference writes ``pattern[x, y] = I`` and blurs ``pattern.T``, i.e. the x/y convention of
        the new api for square shapes (non-square shapes index out of bounds there)."""
        if self.calibration is None:
            raise Exception("Pixel calibration is not set!")
        if shape[0] != shape[1]:
            # (the reference indexes pattern[x, y] on a `shape` array and transposes: out of bounds / transposed output
            # for non-square shapes)
            raise NotImplementedError("non-square shapes are not supported")
        if direct_beam_position is None:
            direct_beam_position = (shape[1] // 2, shape[0] // 2)
        xyz = np.zeros((self.coordinates.shape[0], 3))
        # per-axis calibration (cx, cy), :141-147: the division is done here in float64 exactly as the reference's
        # calibrated_coordinates does it, and K3 runs with a calibration of 1 (x / 1.0 is exact)
        xyz[:, :2] = (self.coordinates[:, :2] + self.offset) / self.calibration
        n = xyz.shape[0]
        dev = engine.device()
        cap = max(32, (n + 31) // 32 * 32)
        X = np.zeros((1, cap, 3))
        X[0, :n] = xyz
        I = np.zeros((1, cap))
        I[0, :n] = self.intensities
        img = engine.render(torch.tensor([n], dtype=torch.int32, device=dev), torch.as_tensor(X, device=dev),
                            torch.as_tensor(I, device=dev), shape, sigma, 1.0,
                            direct_beam_position, in_plane_angle, mirrored, True, True, 1.0)
        return img[0].cpu().numpy().astype(np.float64)
