"""``Simulation2D`` -- result container of the kinematical simulation, B200-native.

Mirrors diffsims/simulations/simulation2d.py:109-557 (constructor contract, ``irot`` / ``iphase``
slicing, iteration, ``get_simulation``, ``polar_flatten_simulations``, ``get_diffraction_pattern``).
Plotting (:559-764) is matplotlib UI and out of scope.

The reference stores an object array of ``DiffractingVector`` (each with a deep-copied, rotated
``Phase``); at 300k-1M templates building those objects costs far more than the GPU work, so results
produced by ``SimulationGenerator`` stay packed on the device (``engine.SpotTable``) and
``DiffractingVector`` objects are materialised lazily by ``coordinates[i]`` / iteration.  Rendering runs
the rasterise kernel (K3) -- one template for the drop-in ``get_diffraction_pattern``, the whole
rotation list for the batched ``get_diffraction_patterns``.
"""
from __future__ import annotations

import copy
from typing import Sequence, Tuple

import numpy as np
import torch

from .. import engine
from ..crystal import Rotation
from ..crystallography import DiffractingVector

__all__ = ["Simulation2D", "get_closest"]


def _is_rotation(obj):
    return hasattr(obj, "to_matrix") and hasattr(obj, "size")


def _is_phase(obj):
    return hasattr(obj, "structure") and not isinstance(obj, (list, tuple, np.ndarray))


class PackedVectors:
    """Sequence of ``DiffractingVector`` backed by one phase's packed device result.

    Behaves like the 1-D object ndarray the reference keeps (``len``, ``size``, ``shape``, integer /
    slice / array indexing, iteration); ``DiffractingVector`` objects are created on access only.
    """

    def __init__(self, phase, rotations, spots, hkl_table, index=None):
        self.phase = phase
        self.rotations = rotations            # Rotation with one entry per row of ``spots``
        self.spots = spots                    # engine.SpotTable (device)
        self.hkl_table = hkl_table            # [n_g, 3] int64 host
        self.index = np.arange(spots.n_rot) if index is None else np.asarray(index)
        self._host = None

    # -- host mirror (one D2H of the packed arrays, on first object access) -----------------
    def host(self):
        if self._host is None:
            s = self.spots
            self._host = dict(count=s.count.cpu().numpy(), g_index=s.g_index.cpu().numpy(),
                              xyz=s.xyz.cpu().numpy(), intensity=s.intensity.cpu().numpy())
        return self._host

    @property
    def size(self):
        return self.index.size

    @property
    def shape(self):
        return (self.index.size,)

    @property
    def ndim(self):
        return 1

    def __len__(self):
        return self.index.size

    def _materialise(self, row):
        h = self.host()
        n = int(h["count"][row])
        G = np.asarray(self.rotations[int(row)].to_matrix()).reshape(3, 3)
        base_phase = self.phase

        def rotated_phase():  # what rotate_with_basis attaches (_diffracting_vector.py:154-158)
            ph = base_phase.deepcopy()
            lat = ph.structure.lattice
            lat.setLatPar(baserot=np.asarray(lat.baserot) @ G)
            return ph

        dv = DiffractingVector(rotated_phase, xyz=h["xyz"][row, :n].copy(),
                               intensity=h["intensity"][row, :n].copy())
        dv._hkl_exact = self.hkl_table[h["g_index"][row, :n]]
        return dv

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self._materialise(self.index[key])
        return PackedVectors(self.phase, self.rotations, self.spots, self.hkl_table, self.index[key])

    def __iter__(self):
        for row in self.index:
            yield self._materialise(row)

    def __array__(self, dtype=None, copy=None):
        out = np.empty(self.size, dtype=object)
        for i, row in enumerate(self.index):
            out[i] = self._materialise(row)
        return out

    # -- device views for the batched kernels ----------------------------------------------
    def device_rows(self):
        s = self.spots
        if self.index.size == s.n_rot and np.array_equal(self.index, np.arange(s.n_rot)):
            return s.count, s.xyz, s.intensity
        idx = torch.as_tensor(self.index, device=s.count.device, dtype=torch.long)
        return s.count[idx].contiguous(), s.xyz[idx].contiguous(), s.intensity[idx].contiguous()


class PhaseGetter:
    """``sim.iphase[...]`` (simulation2d.py:44-75)."""

    def __init__(self, simulation):
        self.simulation = simulation

    def __getitem__(self, item):
        all_phases = self.simulation.phases
        if _is_phase(all_phases):
            raise ValueError("Only one phase in the simulation")
        elif isinstance(item, str):
            ind = [phase.name for phase in all_phases].index(item)
        elif isinstance(item, (int, slice)):
            ind = item
        else:
            raise ValueError("Item must be a string or integer")
        return Simulation2D(
            phases=all_phases[ind],
            coordinates=self.simulation.coordinates[ind],
            rotations=self.simulation.rotations[ind],
            simulation_generator=self.simulation.simulation_generator,
        )


class RotationGetter:
    """``sim.irot[...]`` (simulation2d.py:78-106)."""

    def __init__(self, simulation):
        self.simulation = simulation

    def __getitem__(self, item):
        sim = self.simulation
        if sim.current_size == 1:
            raise ValueError("Only one rotation in the simulation")
        elif _is_phase(sim.phases):
            coords = sim.coordinates[item]
            rotations = sim.rotations[item]
        else:
            coords = [c[item] for c in sim.coordinates]
            rotations = [rot[item] for rot in sim.rotations]
        return Simulation2D(phases=sim.phases, coordinates=coords, rotations=rotations,
                            simulation_generator=sim.simulation_generator)


class Simulation2D:
    """Holds the result of a kinematic diffraction simulation for some phase(s) and rotation(s)."""

    def __init__(self, phases, coordinates, rotations, simulation_generator, reciprocal_radius=1.0):
        if _is_rotation(rotations) and rotations.size == 1:
            if isinstance(coordinates, PackedVectors) and coordinates.size == 1:
                coordinates = coordinates[0]
            if not isinstance(coordinates, DiffractingVector):
                raise ValueError(
                    "If there is only one rotation, then the coordinates must be a DiffractingVector object")
        elif _is_rotation(rotations):
            if not isinstance(coordinates, PackedVectors):
                coordinates = _object_array(coordinates)
            if coordinates.size != rotations.size:
                raise ValueError(
                    f"The number of rotations: {rotations.size} must match the number of "
                    f"coordinates {coordinates.size}")
        else:  # iterable of Rotation, one per phase
            rotations = _object_array(rotations)
            coordinates = _object_array(
                [c if isinstance(c, (PackedVectors, DiffractingVector)) else _object_array(c)
                 for c in coordinates])
            phases = _object_array(phases)
            if rotations.size != phases.size:
                raise ValueError(
                    f"The number of rotations: {rotations.size} must match the number of "
                    f"phases {phases.size}")
            if coordinates.size != phases.size:
                raise ValueError(
                    f"The number of coordinate lists: {coordinates.size} must match the number of "
                    f"phases {phases.size}")
            for r, c in zip(rotations, coordinates):
                n_c = 1 if isinstance(c, DiffractingVector) else len(c)
                if r.size != n_c:
                    raise ValueError(
                        f"The number of rotations: {r.size} must match the number of coordinates {n_c}")
        self.phases = phases
        self.rotations = rotations
        self.coordinates = coordinates
        self.simulation_generator = simulation_generator

        self.phase_index = 0
        self.rotation_index = 0
        self._rot_plot = None
        self._diff_plot = None
        self.reciporical_radius = reciprocal_radius  # (sic) attribute name of the reference

        self.iphase = PhaseGetter(self)
        self.irot = RotationGetter(self)
        self._rotation_slider = None
        self._phase_slider = None

    # ------------------------------------------------------------------ access
    def get_simulation(self, item):
        """Return the rotation, the phase index and the coordinates of flat index ``item`` (:200-218)."""
        if self.has_multiple_phases:
            cumsum = np.cumsum(self._num_rotations())
            ind = np.searchsorted(cumsum, item, side="right")
            cumsum = np.insert(cumsum, 0, 0)
            num_rot = cumsum[ind]
            if self.has_multiple_rotations[ind]:
                return (self.rotations[ind][item - num_rot], ind, self.coordinates[ind][item - num_rot])
            else:
                return self.rotations[ind], ind, self.coordinates[ind]
        elif self.has_multiple_rotations:
            return self.rotations[item], 0, self.coordinates[item]
        else:
            return self.rotations[item], 0, self.coordinates

    def _num_rotations(self):
        if self.has_multiple_phases:
            return [r.size for r in self.rotations]
        else:
            return self.rotations.size

    def __iter__(self):
        return self

    def __next__(self):
        if self.phase_index == self.num_phases:
            self.phase_index = 0
            raise StopIteration
        coords = self.coordinates[self.phase_index] if self.has_multiple_phases else self.coordinates
        multi = self.has_multiple_rotations
        if self.has_multiple_phases:
            multi = multi[self.phase_index]
        if multi:
            coords = coords[self.rotation_index]
        if self.rotation_index + 1 == self.current_size:
            self.rotation_index = 0
            self.phase_index += 1
        else:
            self.rotation_index += 1
        return coords

    @property
    def current_size(self):
        """Number of rotations in the current phase."""
        if self.has_multiple_phases:
            return self.rotations[self.phase_index].size
        return self.rotations.size

    def deepcopy(self):
        return copy.deepcopy(self)

    @property
    def current_phase(self):
        return self.phases[self.phase_index] if self.has_multiple_phases else self.phases

    @property
    def num_phases(self):
        if hasattr(self.phases, "__len__"):
            return len(self.phases)
        return 1

    @property
    def has_multiple_phases(self):
        return self.num_phases > 1

    @property
    def has_multiple_rotations(self):
        if _is_rotation(self.rotations):
            return self.rotations.size > 1
        return [r.size > 1 for r in self.rotations]

    def get_current_coordinates(self):
        """(A copy of) the DiffractingVector of the current phase and rotation (:465-474)."""
        if self.has_multiple_phases:
            c = self.coordinates[self.phase_index]
            if not isinstance(c, DiffractingVector):
                c = c[self.rotation_index]
            return copy.deepcopy(c)
        elif self.has_multiple_rotations:
            return copy.deepcopy(self.coordinates[self.rotation_index])
        return copy.deepcopy(self.coordinates)

    def get_current_rotation_matrix(self):
        if self.has_multiple_phases:
            return copy.deepcopy(self.rotations[self.phase_index].to_matrix()[self.rotation_index])
        return copy.deepcopy(self.rotations.to_matrix()[self.rotation_index])

    # ------------------------------------------------------------------ coordinates
    def _get_transformed_coordinates(self, angle, center=(0, 0), mirrored=False, units="real",
                                     calibration=None):
        """Translate, rotate or mirror the spot coordinates of the current pattern (:261-285)."""
        coords = self.get_current_coordinates()
        if units != "real":
            center = np.array(center)
            coords.data[...] = coords.data / calibration
        cx, cy = center
        x = coords.data[:, 0].copy()
        y = coords.data[:, 1].copy()
        mirrored_factor = -1 if mirrored else 1
        theta = mirrored_factor * np.arctan2(y, x) + np.deg2rad(angle)
        rd = np.sqrt(x ** 2 + y ** 2)
        coords[:, 0] = rd * np.cos(theta) + cx
        coords[:, 1] = rd * np.sin(theta) + cy
        return coords

    def rotate_shift_coordinates(self, angle, center=(0, 0), mirrored=False):
        """Rotate, flip or shift patterns in-plane (:294-311)."""
        return self._get_transformed_coordinates(angle, center, mirrored, units="real")

    def polar_flatten_simulations(self, radial_axes=None, azimuthal_axes=None):
        """(n_simulations, max_spots) arrays of r, theta, intensity for template matching (:313-355)."""
        packed = self._packed_phases()
        if packed is not None and self.phase_index == 0 and self.rotation_index == 0:
            # packed device result: one kernel per phase instead of a Python loop over templates
            rows = [p.device_rows() for p in packed]
            max_num_spots = max(int(c.max().item()) if c.numel() else 0 for c, _, _ in rows)
            outs = [engine.polar_flatten(c, x, i, max_num_spots, radial_axes, azimuthal_axes) for c, x, i in rows]
            r, t, inten = (torch.cat([o[k] for o in outs]).cpu().numpy() for k in range(3))
            if radial_axes is not None and azimuthal_axes is not None:
                r, t = r.astype(int), t.astype(int)
            return r, t, inten
        # plain DiffractingVector objects (a user-built Simulation2D): same kernel on packed copies; the
        # iteration is stateful like the reference's (it starts at the current phase / rotation index)
        flattened_vectors = [sim for sim in self]
        max_num_spots = max([v.size for v in flattened_vectors])
        count, xyz, inten = pack_vectors(flattened_vectors, engine.device())
        r, t, i = (o.cpu().numpy() for o in engine.polar_flatten(count, xyz, inten, max_num_spots, radial_axes,
                                                                 azimuthal_axes))
        if radial_axes is not None and azimuthal_axes is not None:
            r, t = r.astype(int), t.astype(int)
        return r, t, i

    def _packed_phases(self):
        """The per-phase PackedVectors when the whole result is still packed on the device, else None."""
        if isinstance(self.coordinates, PackedVectors):
            return [self.coordinates]
        if isinstance(self.coordinates, np.ndarray) and self.coordinates.dtype == object and len(self.coordinates) \
                and all(isinstance(c, PackedVectors) for c in self.coordinates):
            return list(self.coordinates)
        return None

    # ------------------------------------------------------------------ rendering
    def get_diffraction_pattern(self, shape: Tuple[int, int] = None, sigma: float = 10,
                                direct_beam_position: Tuple[int, int] = None, in_plane_angle: float = 0,
                                calibration: float = 0.01, mirrored: bool = False, fast: bool = True,
                                normalize: bool = True, clip_threshold: float = 1):
        """Diffraction pattern of the current (phase, rotation) as a numpy array with a 2-D Gaussian
        per reflection (:357-442).  The rasterisation runs on the GPU (K3); the result is float64 like
        the reference's, computed in float32 (within 1e-4 of the peak)."""
        coords = self.get_current_coordinates()
        is_int = np.issubdtype(np.asarray(coords.data).dtype, np.integer)
        if direct_beam_position is None:
            if fast or is_int:
                direct_beam_position = (shape[1] // 2, shape[0] // 2)
            else:
                direct_beam_position = ((shape[1] - 1) / 2, (shape[0] - 1) / 2)
        xyz = np.asarray(coords.data)
        if is_int:
            # integer-dtype vectors keep an integer array in the reference: every assignment truncates,
            # and the integer rasteriser branch is taken whatever ``fast`` says
            # (detector_functions.py:293; tests/simulations/test_simulations2d.py:133-166)
            t = self._get_transformed_coordinates(in_plane_angle, direct_beam_position, mirrored,
                                                  units="pixel", calibration=calibration)
            xyz = np.asarray(t.data, dtype=float)
            calibration, direct_beam_position, in_plane_angle, mirrored, fast = 1.0, (0, 0), 0.0, False, True
        inten = np.asarray(coords.intensity, dtype=float)
        n = inten.shape[0]
        dev = engine.device()
        cap = max(32, (n + 31) // 32 * 32)
        X = np.zeros((1, cap, 3))
        X[0, :n] = np.asarray(xyz, dtype=float).reshape(-1, 3)
        I = np.zeros((1, cap))
        I[0, :n] = inten
        img = engine.render(torch.tensor([n], dtype=torch.int32, device=dev), torch.as_tensor(X, device=dev),
                            torch.as_tensor(I, device=dev), shape, sigma, calibration, direct_beam_position,
                            in_plane_angle, mirrored, fast, normalize, clip_threshold)
        return img[0].cpu().numpy().astype(np.float64)

    def get_diffraction_patterns(self, shape: Tuple[int, int], sigma: float = 10,
                                 direct_beam_position: Tuple[int, int] = None, in_plane_angle: float = 0,
                                 calibration: float = 0.01, mirrored: bool = False, fast: bool = True,
                                 normalize: bool = True, clip_threshold: float = 1, out=None):
        """Batched ``get_diffraction_pattern``: every rotation of the (single or current) phase in one
        kernel launch.  Returns a float32 device tensor [n_rotations, H, W] (or fills ``out``).  This
        is the B200-native extension the reference lacks (it renders one template per call)."""
        coords = self.coordinates[self.phase_index] if self.has_multiple_phases else self.coordinates
        if direct_beam_position is None:
            if fast:
                direct_beam_position = (shape[1] // 2, shape[0] // 2)
            else:
                direct_beam_position = ((shape[1] - 1) / 2, (shape[0] - 1) / 2)
        if isinstance(coords, PackedVectors):
            count, xyz, inten = coords.device_rows()
        else:
            vecs = [coords] if isinstance(coords, DiffractingVector) else list(coords)
            count, xyz, inten = pack_vectors(vecs, engine.device())
        return engine.render(count, xyz, inten, shape, sigma, calibration, direct_beam_position,
                             in_plane_angle, mirrored, fast, normalize, clip_threshold, out=out)

    def plot(self, *args, **kwargs):  # pragma: no cover
        raise NotImplementedError("plotting is matplotlib UI and out of scope of diffsims_b200")

    plot_rotations = plot


def pack_vectors(vectors: Sequence[DiffractingVector], dev):
    """Pack host DiffractingVectors into the padded device layout K3 reads."""
    n = np.array([v.size for v in vectors], dtype=np.int32)
    cap = max(32, (int(n.max()) + 31) // 32 * 32) if len(n) else 32
    X = np.zeros((len(vectors), cap, 3))
    I = np.zeros((len(vectors), cap))
    for i, v in enumerate(vectors):
        X[i, : n[i]] = np.asarray(v.data, dtype=float)
        I[i, : n[i]] = np.asarray(v.intensity, dtype=float)
    return torch.as_tensor(n, device=dev), torch.as_tensor(X, device=dev), torch.as_tensor(I, device=dev)


def _object_array(seq):
    if isinstance(seq, np.ndarray) and seq.dtype == object:
        return seq
    seq = list(seq) if not isinstance(seq, (DiffractingVector, PackedVectors)) else [seq]
    out = np.empty(len(seq), dtype=object)
    for i, v in enumerate(seq):
        out[i] = v
    return out


def get_closest(array, values):
    """Index of the closest entry of sorted ``array`` for each value (:767-781)."""
    array = np.array(array)
    idxs = np.searchsorted(array, values, side="left")
    prev_idx_is_less = (idxs == len(array)) | (
        np.fabs(values - array[np.maximum(idxs - 1, 0)])
        < np.fabs(values - array[np.minimum(idxs, len(array) - 1)]))
    idxs[prev_idx_is_less] -= 1
    return idxs
