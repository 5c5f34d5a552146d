"""Vertices of the sphere meshes ``get_beam_directions_grid`` can be built on.

Same names, arguments and return values as diffsims/generators/sphere_mesh_generators.py.  These are short
host-side vertex lists (a few 1e5 rows at 0.5 degrees); the crop to the stereographic triangle, the ordered
compaction and the conversion to Euler angles / quaternions run on the device (``ds_beam_grid`` for the cube
meshes, which never materialise their vertices on the host, ``ds_beam_points`` for the meshes below).
"""
import numpy as np
from scipy.spatial import cKDTree

from .. import engine

__all__ = ["beam_directions_grid_to_euler", "get_cube_mesh_vertices", "get_icosahedral_mesh_vertices",
           "get_random_sphere_vertices", "get_uv_sphere_mesh_vertices"]


def _unit(v):
    return (v.T / np.linalg.norm(v, axis=1)).T


def get_uv_sphere_mesh_vertices(resolution):
    """Latitude / longitude mesh, one vertex per pole (reference :42-93): azimuth psi is the slow index,
    elevation theta the fast one; x = cos(psi) sin(theta), y = sin(psi) sin(theta), z = cos(theta)."""
    n_theta = int(np.ceil(180 / resolution)) + 1
    n_psi = int(np.ceil(360 / resolution))
    psi = np.repeat(np.linspace(0, 2 * np.pi, num=n_psi, endpoint=False), n_theta)
    theta = np.tile(np.linspace(0, np.pi, num=n_theta, endpoint=True), n_psi)
    pole = (theta == 0) | ~(theta < np.deg2rad(180))
    keep = ~pole | (psi == 0)
    psi, theta = psi[keep], theta[keep]
    one = np.ones(psi.shape[0])
    return np.stack([one * np.cos(psi) * np.sin(theta), one * np.sin(psi) * np.sin(theta), one * np.cos(theta)],
                    axis=1)


def get_cube_mesh_vertices(resolution, grid_type="spherified_corner"):
    """Cube mesh projected on the sphere (reference :96-197): bottom, top, east, west, south, north faces of the
    tan-spaced face grid, then the two corners the half-open grids miss."""
    from .rotation_list_generators import _face_grid
    i = _face_grid(resolution, grid_type)
    y, x = (a.ravel() for a in np.meshgrid(i, i, indexing="ij"))
    z = np.ones_like(x)
    faces = [(-x, -y, -z), (x, y, z), (z, x, -y), (-z, -x, y), (x, -z, y), (-x, z, -y)]
    pts = np.concatenate([np.stack(f, axis=1) for f in faces] + [np.array([[-1.0, 1, 1], [1, -1, -1]])])
    return _unit(pts)


_T = (1.0 + np.sqrt(5.0)) / 2.0
_ICOSAHEDRON = np.array([(-1, _T, 0), (1, _T, 0), (-1, -_T, 0), (1, -_T, 0), (0, -1, _T), (0, 1, _T),
                         (0, -1, -_T), (0, 1, -_T), (_T, 0, -1), (_T, 0, 1), (-_T, 0, -1), (-_T, 0, 1)])
_ICOSAHEDRON_FACES = ((0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
                      (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
                      (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1))


def _edge_order(faces):
    """Edges in the order the reference visits them (:228-234): it gathers the sorted vertex pairs of every
    face into a ``set`` and iterates that, so the same insertions are replayed here."""
    seen = set()
    for a, b, c in faces:
        for e in ((a, b), (b, c), (c, a)):
            seen.add((min(e), max(e)))
    return list(seen)


def _refined_vertices(corners, faces, n):
    """Vertices of the n-fold subdivision (reference `_compose_from_faces` :200-322, which also builds a cell
    table it never returns): corners, n - 1 nodes on every edge, then the interior nodes of every face
    (barycentric grid, row i = 1 .. n - 1, column j = 1 .. n - i - 1)."""
    blocks = [corners]
    t = np.linspace(1 / n, 1.0, n - 1, endpoint=False)
    for i0, i1 in _edge_order(faces):
        blocks.append(np.outer(1 - t, corners[i0]) + np.outer(t, corners[i1]))
    if n > 1:
        rows = np.concatenate([np.full(n - i - 1, i) for i in range(1, n)]) / n
        cols = np.concatenate([np.arange(1, n - i) for i in range(1, n)]) / n
        bary = np.array([1.0 - rows - cols, cols, rows])
        for f in faces:
            blocks.append(np.dot(corners[list(f)].T, bary).T)
    return np.concatenate(blocks)


def _max_neighbour_angle(vertices, leaf_size=50):
    """Largest angle (degrees) between a vertex and its nearest neighbour (reference :325-375)."""
    v = _unit(vertices)
    nearest = cKDTree(v, leaf_size).query(v, k=2)[1][:, 1]
    return np.max(np.rad2deg(np.arccos(np.sum(v * v[nearest], axis=1))))


def get_icosahedral_mesh_vertices(resolution):
    """Icosahedron refined until neighbouring vertices are at most ``resolution`` degrees apart (reference
    :378-450)."""
    n, angle, vertices = 1, _max_neighbour_angle(_ICOSAHEDRON), None
    while angle > resolution:
        vertices = _refined_vertices(_ICOSAHEDRON, _ICOSAHEDRON_FACES, n)
        angle = _max_neighbour_angle(vertices)
        n += 1
    if vertices is None:
        raise ValueError(f"resolution {resolution} is coarser than the icosahedron itself")
    return (vertices.T / np.sqrt(np.einsum("ij,ij->i", vertices, vertices)).T).T


def get_random_sphere_vertices(resolution, seed=None):
    """Normally distributed points pushed to the sphere (reference :453-483); ``resolution`` is the expected
    mean nearest-neighbour angle."""
    number = int(1 / (4 * np.pi) * (360 / resolution) ** 2)
    rng = np.random.default_rng() if seed is None else np.random.default_rng(seed=seed)
    return _unit(rng.normal(size=(number, 3)))


def beam_directions_grid_to_euler(vectors):
    """Euler angles (phi1 = 0, Phi, phi2) in degrees bringing z onto each vector (reference :486-526), computed
    by the ``ds_beam_points`` kernel."""
    from .rotation_list_generators import _points_to_grid
    euler, _ = _points_to_grid(np.asarray(vectors, dtype=np.float64), "triclinic", want_quaternions=False)
    return euler.cpu().numpy()
