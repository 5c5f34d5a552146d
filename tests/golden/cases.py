"""Seeded inputs shared by make_golden.py (which runs the reference) and the tests.

Structures are plain numbers; they are built with the duck-typed stand-ins from
``diffsims_b200.crystal`` so neither orix nor diffpy is needed.
"""
import numpy as np

from diffsims_b200.crystal import Atom, Lattice, Phase, Structure

DW = {"Si": 0.5, "O": 0.7, "Fe": 0.35, "C": 0.45}


def _pnma_8d(x, y, z):
    return [(x, y, z), (-x + .5, -y, z + .5), (-x, y + .5, -z), (x + .5, -y + .5, -z + .5),
            (-x, -y, -z), (x + .5, y, -z + .5), (x, -y + .5, z), (-x + .5, y + .5, z + .5)]


def _pnma_4c(x, z):
    return [(x, .25, z), (-x + .5, .75, z + .5), (-x, .75, -z), (x + .5, .25, -z + .5)]


def _si_atoms():
    out = []
    for c in [[0, 0, 0], [0.5, 0, 0.5], [0, 0.5, 0.5], [0.5, 0.5, 0]]:
        out.append(("Si", c, 1.0))
        out.append(("Si", [c[0] + 0.25, c[1] + 0.25, c[2] + 0.25], 1.0))
    return out


def _large_atoms(n=500, seed=1):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(size=(n, 3))
    els = ["O", "Si", "Al", "Ca", "Fe"]
    return [(els[i % 5], xyz[i], 1.0) for i in range(n)]


STRUCTURES = {
    # name: (lattice params, [(element, frac xyz, occupancy)], space group, point group)
    "si": ((5.431, 5.431, 5.431, 90, 90, 90), _si_atoms(), 227, "m-3m"),
    "al": ((4.05, 4.05, 4.05, 90, 90, 90),
           [("Al", p, 1.0) for p in ([0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5])], 225, "m-3m"),
    "graphite": ((2.464, 2.464, 6.711, 90, 90, 120),
                 [("C", [0, 0, .25], 1.0), ("C", [0, 0, .75], 1.0),
                  ("C", [1 / 3, 2 / 3, .25], 1.0), ("C", [2 / 3, 1 / 3, .75], 1.0)], 194, "6/mmm"),
    "ti": ((2.95, 2.95, 4.68, 90, 90, 120),
           [("Ti", [1 / 3, 2 / 3, .25], 1.0), ("Ti", [2 / 3, 1 / 3, .75], 1.0)], 194, "6/mmm"),
    "fe_bcc": ((2.8665, 2.8665, 2.8665, 90, 90, 90),
               [("Fe", [0, 0, 0], 1.0), ("Fe", [.5, .5, .5], 1.0)], 229, "m-3m"),
    "fe_fcc": ((3.59, 3.59, 3.59, 90, 90, 90),
               [("Fe", p, 1.0) for p in ([0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5])], 225, "m-3m"),
    "fe3c": ((5.09, 6.74, 4.53, 90, 90, 90),
             [("Fe", p, 1.0) for p in _pnma_8d(0.186, 0.063, 0.328)]
             + [("Fe", p, 1.0) for p in _pnma_4c(0.036, 0.852)]
             + [("C", p, 1.0) for p in _pnma_4c(0.890, 0.450)], 62, "mmm"),
    # oblique cell with partial occupancy and an ionic label: exercises the realignment,
    # get_element() and the occupancy path
    "triclinic": ((4.1, 5.3, 6.2, 81.0, 97.0, 112.0),
                  [("Si4+", [0.1, 0.2, 0.3], 1.0), ("O2-", [0.6, 0.15, 0.85], 0.5),
                   ("Fe", [0.33, 0.71, 0.42], 0.75), ("O", [0.9, 0.5, 0.05], 1.0)], 1, "-1"),
}


def structure(name):
    if name == "large":
        lat = Lattice(12.0, 12.0, 12.0, 90, 90, 90)
        return Structure([Atom(e, x, o) for e, x, o in _large_atoms()], lat)
    p, atoms, _, _ = STRUCTURES[name]
    return Structure([Atom(e, x, o) for e, x, o in atoms], Lattice(*p))


def phase(name):
    if name == "large":
        return Phase("large", space_group=1, structure=structure(name))
    _, _, sg, pg = STRUCTURES[name]
    return Phase(name, space_group=sg, point_group=pg, structure=structure(name))


def hkl_box(n):
    r = np.arange(-n, n + 1)
    return np.array(np.meshgrid(r, r, r, indexing="ij")).reshape(3, -1).T.astype(float)


# name: (shape (H, W), sigma, n_spots, seed)
DETECTOR_CASES = {
    "a": ((64, 64), 2.0, 12, 0),
    "b": ((144, 144), 10.0, 25, 1),   # sigma > distance to border: reflect folding
    "c": ((50, 80), 3.5, 40, 2),      # non-square, duplicate pixels likely
    "d": ((32, 32), 10.0, 6, 3),      # kernel radius (40) larger than the image
}


def detector_spots(shape, n, seed):
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(0, shape[1], n), rng.uniform(0, shape[0], n)], axis=1)
    xy[n // 2] = xy[0]  # a duplicate pixel: the integer branch is last-write-wins
    xy[1] = (0.2, 0.7)  # corner
    inten = rng.uniform(0.1, 5.0, n)
    return xy, inten


def detector_spots_outside(shape, n, seed):
    """Float pixel coordinates partly outside the frame (the bare rasteriser spreads them into it; a negative
    slice stop wraps around as in numpy) and integer coordinates with negative entries (index wrap)."""
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(-8, shape[1] + 8, n), rng.uniform(-8, shape[0] + 8, n)], axis=1)
    xy_int = np.stack([rng.integers(-shape[1], shape[1], n), rng.integers(-shape[0], shape[0], n)], axis=1)
    return xy, xy_int, rng.uniform(20, 900, n)


def random_eulers(n, seed):
    rng = np.random.default_rng(seed)
    e = np.stack([rng.uniform(0, 360, n), np.rad2deg(np.arccos(rng.uniform(-1, 1, n))),
                  rng.uniform(0, 360, n)], axis=1)
    return [tuple(float(v) for v in row) for row in e]


ED_CASES = {
    "si": dict(structure="si", kv=200, rr=1.0, s_max=0.01, with_direct_beam=True,
               calibration=1.0 / 128, half_shape=(128, 128), shape=(256, 256), sigma=10,
               eulers=[(0, 0, 0), (0, 45, 0)] + random_eulers(6, 0)),
    "graphite": dict(structure="graphite", kv=200, rr=1.6768, s_max=0.1, with_direct_beam=False,
                     calibration=0.0262, half_shape=(64, 64), shape=(128, 128), sigma=1.4,
                     eulers=[(0, 90, 120), (10, 20, 30)] + random_eulers(4, 1)),
    "fe3c_linear": dict(structure="fe3c", kv=300, rr=1.2, s_max=0.03, with_direct_beam=True,
                        model="linear", scattering_params="xtables", dw=DW,
                        calibration=0.01, half_shape=(72, 72), shape=(144, 144), sigma=4,
                        eulers=[(0, 0, 0)] + random_eulers(5, 2)),
    "triclinic_sinc": dict(structure="triclinic", kv=120, rr=0.9, s_max=0.02,
                           with_direct_beam=True, model="sinc", minimum_intensity=1e-6,
                           calibration=0.008, half_shape=(72, 72), shape=(144, 144), sigma=2,
                           eulers=random_eulers(4, 3)),
}


BEAM_GRID_SYSTEMS = ("cubic", "hexagonal", "trigonal", "tetragonal", "orthorhombic", "monoclinic", "triclinic")
# diffsims/tests/generators/test_rotation_list_generator.py:80-94
BEAM_GRID_SIZES_2DEG = dict(cubic=300, hexagonal=1050, trigonal=1657, tetragonal=852, orthorhombic=1657,
                            monoclinic=6441, triclinic=12698)
