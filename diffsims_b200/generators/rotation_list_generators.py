"""Rotation-list producers next to the hot path, B200-native.

``get_beam_directions_grid`` keeps the reference's signature and return type
(diffsims/generators/rotation_list_generators.py:176-267: an [N, 3] array of Bunge Euler angles in
degrees with phi1 = 0); the mesh points, the crop to the stereographic triangle and the conversion run
in the ``ds_beam_grid`` kernel.  ``beam_directions_device`` is the same grid left in HBM -- Euler
angles and the active quaternions the simulate kernel consumes -- so that a 3e5 - 1e6 entry rotation list
never exists as a Python list of tuples.

All six meshes of the reference are available.  The cube meshes (``normalized_cube``, ``spherified_cube_edge``
-- the default --, ``spherified_cube_corner``) are generated inside the kernel from the 1-D face grid; the
uv-sphere / icosahedral / random vertex lists come from ``sphere_mesh_generators`` and are cropped and converted
by ``ds_beam_points``.  The orix-backed fundamental-zone and local grids (``get_fundamental_zone_grid``,
``get_local_grid``: orix ``get_sample_fundamental`` / ``get_sample_local``) are not built.
"""
import math

import numpy as np
import torch

from .. import _cabi, engine
from ..crystal import Rotation

__all__ = ["get_beam_directions_grid", "beam_directions_device", "get_grid_around_beam_direction",
           "crystal_system_dictionary"]

# triangle corners per crystal system (rotation_list_generators.py:42-55)
crystal_system_dictionary = {
    "cubic": [(0, 0, 1), (1, 1, 1), (1, 0, 1)],
    "hexagonal": [(0, 0, 0, 1), (9, 1, -10, 0), (2, -1, -1, 0)],
    "trigonal": [(0, 0, 0, 1), (-2, 1, 1, 0), (-1, 2, -1, 0)],
    "tetragonal": [(0, 0, 1), (1, 0, 0), (1, 1, 0)],
    "orthorhombic": [(0, 0, 1), (-1, 0, 0), (0, 1, 0)],
    "monoclinic": [(0, -1, 0), (0, 0, 1), (0, 1, 0)],
}


def _uvtw_to_uvw(uvtw):
    # diffsims/utils/sim_utils.py:496-510
    u, v, t, w = uvtw
    u, v, w = 2 * u + v, 2 * v + u, w
    g = math.gcd(math.gcd(u, v), w)
    return tuple(int(x / g) for x in (u, v, w))


def _face_grid(resolution, grid_type):
    """The 1-D grid of a cube face, exactly as get_cube_mesh_vertices builds it
    (sphere_mesh_generators.py:156-183)."""
    max_angle, max_dist = np.deg2rad(45), 1
    if grid_type == "normalized":
        steps = np.ceil(max_dist / np.tan(np.deg2rad(resolution)))
        return np.arange(-steps, steps) / steps
    if grid_type == "spherified_edge":
        steps = np.ceil(np.rad2deg(max_angle) / resolution)
        return np.tan(np.arange(-steps, steps) * (np.arctan(max_dist) / steps))
    if grid_type == "spherified_corner":
        steps = np.ceil(np.arccos(1 / np.sqrt(3)) / np.deg2rad(resolution))
        return np.tan(np.arange(-steps, steps) * (np.arctan(np.sqrt(2)) / steps)) / np.sqrt(2)
    raise ValueError(f"grid type {grid_type} not a valid grid type. "
                     f"Valid options: normalized, spherified_edge, spherified_corner.")


def _crop(crystal_system):
    """(mode, normals[3][3]) of the crop the reference applies (:238-266)."""
    if crystal_system == "triclinic":
        return 0, None
    if crystal_system == "monoclinic":
        # the reference's second filter overwrites its first: only x >= epsilon is applied
        return 1, None
    a, b, c = crystal_system_dictionary[crystal_system]
    if len(a) == 4:
        a, b, c = _uvtw_to_uvw(a), _uvtw_to_uvw(b), _uvtw_to_uvw(c)
    a, b, c = (np.asarray(v, dtype=float) for v in (a, b, c))
    nrm = np.array([np.dot(np.cross(a, b), c) * np.cross(a, b),
                    np.dot(np.cross(b, c), a) * np.cross(b, c),
                    np.dot(np.cross(c, a), b) * np.cross(c, a)], dtype=float)
    return 2, np.ascontiguousarray(nrm)


def _compact(call, n_blocks, dev, want_euler, want_quaternions):
    """Two-pass ordered compaction shared by ds_beam_grid / ds_beam_points: count, scan, fill."""
    counts = torch.empty(n_blocks, dtype=torch.int32, device=dev)
    call(0, counts, None, None, None)
    incl = torch.cumsum(counts, dim=0, dtype=torch.int64)
    offsets = (incl - counts).contiguous()
    n = int(incl[-1].item())
    euler = torch.empty((n, 3), dtype=torch.float64, device=dev) if want_euler else None
    quat = torch.empty((n, 4), dtype=torch.float64, device=dev) if want_quaternions else None
    if n:
        call(1, counts, offsets, euler, quat)
    return euler, quat


def _points_to_grid(points, crystal_system, want_euler=True, want_quaternions=True):
    """Crop mesh vertices [N, 3] to the triangle of ``crystal_system`` and convert them on the device."""
    if crystal_system not in crystal_system_dictionary and crystal_system != "triclinic":
        raise KeyError(crystal_system)
    dev = engine.device()
    pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64), device=dev)
    mode, nrm = _crop(crystal_system)
    nrm_p = None if nrm is None else nrm.ctypes.data_as(_cabi.c_void_p)
    lib = _cabi.lib()

    def call(pass_, counts, offsets, euler, quat):
        _cabi.check(lib.ds_beam_points(engine._stream(), pass_, pts.shape[0], _cabi.ptr(pts), mode, nrm_p, -1e-13,
                                       _cabi.ptr(counts), _cabi.ptr(offsets), _cabi.ptr(euler), _cabi.ptr(quat)),
                    "ds_beam_points")

    if pts.shape[0] == 0:
        empty = lambda k: torch.empty((0, k), dtype=torch.float64, device=dev)  # noqa: E731
        return (empty(3) if want_euler else None), (empty(4) if want_quaternions else None)
    return _compact(call, int(lib.ds_beam_points_num_blocks(pts.shape[0])), dev, want_euler, want_quaternions)


def beam_directions_device(crystal_system, resolution, mesh="spherified_cube_edge", want_euler=True,
                           want_quaternions=True):
    """Beam-direction grid in HBM: returns (euler_deg [N,3] or None, active_quaternions [N,4] or None)."""
    from . import sphere_mesh_generators as smg
    if mesh == "uv_sphere":
        return _points_to_grid(smg.get_uv_sphere_mesh_vertices(resolution), crystal_system, want_euler,
                               want_quaternions)
    if mesh == "icosahedral":
        return _points_to_grid(smg.get_icosahedral_mesh_vertices(resolution), crystal_system, want_euler,
                               want_quaternions)
    if mesh == "random":
        return _points_to_grid(smg.get_random_sphere_vertices(resolution), crystal_system, want_euler,
                               want_quaternions)
    if mesh == "spherified_cube_corner":
        grid_type = "spherified_corner"
    elif mesh in ("normalized_cube", "spherified_cube_edge"):
        if crystal_system == "hexagonal":  # :216-218
            resolution = resolution / np.sqrt(2)
        grid_type = "normalized" if mesh == "normalized_cube" else "spherified_edge"
    else:
        raise NotImplementedError(
            f"The mesh {mesh} is not recognized. Please use: uv_sphere, normalized_cube, "
            f"spherified_cube_edge, spherified_cube_corner, icosahedral, random")
    if crystal_system not in crystal_system_dictionary and crystal_system != "triclinic":
        raise KeyError(crystal_system)
    dev = engine.device()
    i_vals = torch.as_tensor(np.ascontiguousarray(_face_grid(resolution, grid_type)), device=dev)
    n_i = i_vals.numel()
    mode, nrm = _crop(crystal_system)
    nrm_p = None if nrm is None else nrm.ctypes.data_as(_cabi.c_void_p)
    lib = _cabi.lib()

    def call(pass_, counts, offsets, euler, quat):
        _cabi.check(lib.ds_beam_grid(engine._stream(), pass_, n_i, _cabi.ptr(i_vals), mode, nrm_p, -1e-13,
                                     _cabi.ptr(counts), _cabi.ptr(offsets), _cabi.ptr(euler), _cabi.ptr(quat)),
                    "ds_beam_grid")

    return _compact(call, int(lib.ds_beam_grid_num_blocks(n_i)), dev, want_euler, want_quaternions)


def get_beam_directions_grid(crystal_system, resolution, mesh="spherified_cube_edge"):
    """Array of beam directions within the stereographic triangle of ``crystal_system`` as Euler angles
    (degrees); same signature and return value as the reference (:176-267)."""
    euler, _ = beam_directions_device(crystal_system, resolution, mesh, want_quaternions=False)
    return euler.cpu().numpy()


def get_grid_around_beam_direction(beam_rotation, resolution, angular_range=(0, 360)):
    """Rotations about a beam direction (:137-173): ``beam_rotation * Rz(angle)`` as Euler tuples rounded to
    two decimals (host arithmetic on at most 360 / resolution entries)."""
    beam = Rotation.from_euler(np.deg2rad(np.asarray(beam_rotation, dtype=float)))
    angles = np.deg2rad(np.arange(start=angular_range[0], stop=angular_range[1], step=resolution))
    in_plane = Rotation(np.stack([np.cos(angles / 2), np.zeros_like(angles), np.zeros_like(angles),
                                  np.sin(angles / 2)], axis=1))
    grid = Rotation(np.repeat(beam.data, angles.shape[0], axis=0)) * in_plane
    return [tuple(np.round(np.rad2deg(e), decimals=2)) for e in grid.to_euler().tolist()]
