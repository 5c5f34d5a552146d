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
by ``ds_beam_points``.

``get_fundamental_zone_grid`` / ``get_local_grid`` / ``get_list_from_orix`` (:58-134) call
``orix.sampling.get_sample_fundamental`` / ``get_sample_local`` in the reference.  orix is not under
/root/reference and not installed, so these are native (``ds_so3_grid``: the published cubochoric equal-volume
grid + fundamental-zone / angle crop, see csrc/so3_grid.cu) and **parity with orix's point lists is unpinned**;
what is pinned: the reference's own test (a non-empty list of tuples), the equal-volume property of the map, the
1 / |G| volume fraction of every fundamental zone and its symmetry-reduction invariants (tests/test_so3_grid.py).
``fundamental_zone_device`` / ``local_grid_device`` leave the list in HBM as active quaternions.
"""
import math

import numpy as np
import torch

from .. import _cabi, engine
from ..crystal import Rotation

__all__ = ["get_beam_directions_grid", "beam_directions_device", "get_grid_around_beam_direction",
           "get_fundamental_zone_grid", "get_local_grid", "get_list_from_orix", "fundamental_zone_device",
           "local_grid_device", "crystal_system_dictionary"]

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


# ----------------------------------------------------------------------------------------------------------------
# SO(3) grids: fundamental zone of a proper point group, neighbourhood of a rotation
# ----------------------------------------------------------------------------------------------------------------
def _axis_angle(axis, deg):
    axis = np.asarray(axis, dtype=float)
    axis = axis / np.linalg.norm(axis)
    h = np.deg2rad(deg) / 2
    return np.concatenate([[np.cos(h)], np.sin(h) * axis])


def _qmul(p, q):
    a1, b1, c1, d1 = p
    a2, b2, c2, d2 = q
    return np.array([a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2, a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
                     a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2, a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2])


# generators of the 11 proper point groups (Schoenflies: C1 C2 D2 C4 D4 C3 D3 C6 D6 T O), axes as in orix / ITA:
# the principal axis along z, the secondary two-fold along x
_PROPER_GENERATORS = {
    "1": [],
    "2": [((0, 0, 1), 180)],
    "222": [((0, 0, 1), 180), ((1, 0, 0), 180)],
    "4": [((0, 0, 1), 90)],
    "422": [((0, 0, 1), 90), ((1, 0, 0), 180)],
    "3": [((0, 0, 1), 120)],
    "32": [((0, 0, 1), 120), ((1, 0, 0), 180)],
    "6": [((0, 0, 1), 60)],
    "622": [((0, 0, 1), 60), ((1, 0, 0), 180)],
    "23": [((0, 0, 1), 180), ((1, 0, 0), 180), ((1, 1, 1), 120)],
    "432": [((0, 0, 1), 90), ((1, 0, 0), 180), ((1, 1, 1), 120)],
}
# point group (Hermann-Mauguin, any of the 32) -> its proper subgroup of the same Laue class
_TO_PROPER = {"1": "1", "-1": "1", "2": "2", "m": "2", "2/m": "2", "222": "222", "mm2": "222", "mmm": "222",
              "4": "4", "-4": "4", "4/m": "4", "422": "422", "4mm": "422", "-42m": "422", "4/mmm": "422",
              "3": "3", "-3": "3", "32": "32", "3m": "32", "-3m": "32", "6": "6", "-6": "6", "6/m": "6",
              "622": "622", "6mm": "622", "-6m2": "622", "6/mmm": "622", "23": "23", "m-3": "23",
              "432": "432", "-43m": "432", "m-3m": "432"}


def proper_point_group_quaternions(name):
    """The rotations of a proper point group as unit quaternions [n, 4] (closure of its generators)."""
    ops = [np.array([1.0, 0.0, 0.0, 0.0])]
    gens = [_axis_angle(ax, ang) for ax, ang in _PROPER_GENERATORS[name]]
    grew = True
    while grew:
        grew = False
        for g in gens:
            for o in list(ops):
                q = _qmul(g, o)
                if q[0] < -1e-12 or (abs(q[0]) < 1e-12 and tuple(q[1:]) < (0, 0, 0)):
                    q = -q
                if not any(np.allclose(q, x, atol=1e-9) or np.allclose(-q, x, atol=1e-9) for x in ops):
                    ops.append(q)
                    grew = True
    return np.ascontiguousarray(np.array(ops))


def _proper_group_of_space_group(space_group):
    sg = int(space_group)
    if not 1 <= sg <= 230:
        raise ValueError("space_group must be between 1 and 230")
    for hi, name in ((2, "1"), (15, "2"), (74, "222"), (88, "4"), (142, "422"), (148, "3"), (167, "32"), (176, "6"),
                     (194, "622"), (206, "23"), (230, "432")):
        if sg <= hi:
            return name


def _resolve_proper_group(point_group, space_group):
    if point_group is not None:
        name = getattr(point_group, "name", point_group)
        if name not in _TO_PROPER:
            raise ValueError(f"unknown point group {name!r}")
        return _TO_PROPER[name]
    if space_group is None:
        raise ValueError("get_fundamental_zone_grid needs a point_group or a space_group")
    return _proper_group_of_space_group(space_group)


def resolution_to_semi_edge_steps(resolution):
    """Cubochoric semi-edge steps N for an average misorientation ``resolution`` (degrees) between neighbours:
    the empirical relation of Singh and De Graef (2016), N = round(131.97049 / (resolution - 0.03732))."""
    return max(1, int(np.round(131.97049 / (float(resolution) - 0.03732))))


def _so3_grid(n_steps, mode, sym=None, max_angle=0.0, centre=None, want_euler=True, want_quaternions=True):
    dev = engine.device()
    lib = _cabi.lib()
    sym = None if sym is None else np.ascontiguousarray(sym, dtype=np.float64)
    centre = None if centre is None else np.ascontiguousarray(centre, dtype=np.float64).reshape(4)
    sym_p = None if sym is None else sym.ctypes.data_as(_cabi.c_void_p)
    cen_p = None if centre is None else centre.ctypes.data_as(_cabi.c_void_p)

    def call(pass_, counts, offsets, euler, quat):
        _cabi.check(lib.ds_so3_grid(engine._stream(), pass_, int(n_steps), int(mode), 0 if sym is None else sym.shape[0], sym_p,
                                    float(max_angle), cen_p, _cabi.ptr(counts), _cabi.ptr(offsets), _cabi.ptr(euler),
                                    _cabi.ptr(quat)), "ds_so3_grid")

    return _compact(call, int(lib.ds_so3_grid_num_blocks(int(n_steps))), dev, want_euler, want_quaternions)


def fundamental_zone_device(resolution=2, point_group=None, space_group=None, want_euler=True, want_quaternions=True):
    """Grid of rotations in the fundamental zone, left in HBM: (euler_deg [N, 3] or None, active quaternions [N, 4] or
    None).  ``point_group``: Hermann-Mauguin name (or an object with ``.name``) -- its proper subgroup is used, as orix
    does; else ``space_group`` (1..230)."""
    name = _resolve_proper_group(point_group, space_group)
    return _so3_grid(resolution_to_semi_edge_steps(resolution), 1, sym=proper_point_group_quaternions(name),
                     want_euler=want_euler, want_quaternions=want_quaternions)


def local_grid_device(resolution=2, center=None, grid_width=10, want_euler=True, want_quaternions=True):
    """Grid of rotations within ``grid_width`` degrees of ``center`` (an Euler tuple in degrees, a Rotation or None =
    identity), left in HBM."""
    cq = None
    if center is not None:
        rot = Rotation.from_euler(np.deg2rad(np.asarray(center, dtype=float))) if isinstance(center, (tuple, list)) else center
        cq = np.asarray(rot.data, dtype=float).reshape(-1, 4)[0]
    return _so3_grid(resolution_to_semi_edge_steps(resolution), 2, max_angle=np.deg2rad(grid_width), centre=cq,
                     want_euler=want_euler, want_quaternions=want_quaternions)


def _euler_list(euler, rounding=2):
    return [tuple(row) for row in np.round(euler.cpu().numpy(), decimals=rounding)]


def get_list_from_orix(grid, rounding=2):
    """Converts a grid of rotations (anything with ``to_euler()`` returning radians, as orix's ``Rotation``) to a
    rotation list of Euler tuples in degrees (:58-82)."""
    e = grid.to_euler()
    e = np.asarray(getattr(e, "data", e), dtype=float).reshape(-1, 3)
    return [tuple(np.round(np.rad2deg(row), decimals=rounding)) for row in e]


def get_fundamental_zone_grid(resolution=2, point_group=None, space_group=None):
    """Equispaced grid of rotations within a fundamental zone as a list of Euler tuples (degrees, two decimals); same
    signature and return type as the reference (:85-106).  The reference forwards only ``space_group`` to orix; here
    ``point_group`` is honoured when given."""
    euler, _ = fundamental_zone_device(resolution, point_group, space_group, want_quaternions=False)
    return _euler_list(euler)


def get_local_grid(resolution=2, center=None, grid_width=10):
    """Grid of rotations about a given rotation as a list of Euler tuples (degrees, two decimals); same signature and
    return type as the reference (:109-134)."""
    euler, _ = local_grid_device(resolution, center, grid_width, want_quaternions=False)
    return _euler_list(euler)
