"""Generate the committed golden fixtures by EXECUTING the reference (dev container only).

    python tests/golden/make_golden.py

Writes small ``.npz`` files next to this script.  Sources of truth:

* ``old_simulation.npz``     the reference's own golden image
                             (diffsims/tests/generators/old_simulation.npy, 128x128 f64),
                             re-saved compressed; sha256 of the original recorded inside.
* ``shape_factors.npz``      reference ``utils/shape_factor_models.py`` evaluated on a grid.
* ``detector.npz``           reference ``pattern/detector_functions.py``
                             ``get_pattern_from_pixel_coordinates_and_intensities`` (integer and
                             sub-pixel branches) on seeded random spot lists.
* ``sim_utils.npz``          reference ``utils/sim_utils.py`` ``get_kinematical_intensities``,
                             ``get_points_in_sphere`` on stand-in structures.
* ``beam_grid.npz``          reference ``generators/rotation_list_generators.get_beam_directions_grid``
                             (cube, uv-sphere and icosahedral meshes) for the crystal systems, and the
                             mesh vertices of ``generators/sphere_mesh_generators.py``.
* ``ed_data.npz``            reference OLD api ``DiffractionGenerator.calculate_ed_data`` +
                             ``DiffractionSimulation.get_diffraction_pattern`` (run with the
                             euler2mat placeholder documented in _ref_loader.py).

The structures are rebuilt from plain numbers by tests/golden/cases.py so that the
tests do not depend on this script.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

from _ref_loader import load_reference, REF  # noqa: E402
import cases  # noqa: E402

ns = load_reference()
sfm = ns.shape_factor_models
det = ns.detector_functions
su = ns.sim_utils
dg = ns.diffraction_generator

# ---------------------------------------------------------------- old_simulation
raw = (REF / "diffsims/tests/generators/old_simulation.npy").read_bytes()
old = np.load(REF / "diffsims/tests/generators/old_simulation.npy")
np.savez_compressed(HERE / "old_simulation.npz", image=old,
                    sha256=np.array(hashlib.sha256(raw).hexdigest()))

# ---------------------------------------------------------------- shape factors
s = np.concatenate([np.linspace(-0.12, 0.12, 49), [0.0, 1e-9, -1e-9, 0.01, -0.01, 0.02, 0.004]])
out = {"s": s}
for w in (0.01, 0.05, 0.1):
    out[f"linear_{w}"] = sfm.linear(s.copy(), w)
    out[f"sinc_{w}"] = sfm.sinc(s.copy(), w)
    out[f"sinc7_{w}"] = sfm.sinc(s.copy(), w, minima_number=7)
    out[f"sin2c_{w}"] = sfm.sin2c(s.copy(), w)
    with np.errstate(all="ignore"):
        out[f"atanc_{w}"] = sfm.atanc(s.copy(), w)
    out[f"lorentzian_{w}"] = sfm.lorentzian(s.copy(), w)
    r_spot = np.linspace(0, 2.0, s.size)
    out[f"lorentzian_precession_{w}"] = sfm.lorentzian_precession(s.copy(), w, r_spot, np.deg2rad(0.5))
out["r_spot"] = r_spot
np.savez_compressed(HERE / "shape_factors.npz", **out)

# ---------------------------------------------------------------- detector functions
out = {}
for name, (shape, sigma, n, seed) in cases.DETECTOR_CASES.items():
    xy, inten = cases.detector_spots(shape, n, seed)
    out[f"{name}_int"] = det.get_pattern_from_pixel_coordinates_and_intensities(
        xy.astype(int), inten, shape, sigma)
    out[f"{name}_float"] = det.get_pattern_from_pixel_coordinates_and_intensities(
        xy, inten * 2000.0, shape, sigma, 1.0)
xy, xy_int, inten = cases.detector_spots_outside((70, 90), 40, 5)
out["outside_float"] = det.get_pattern_from_pixel_coordinates_and_intensities(xy, inten, (70, 90), 2.5)
out["outside_int"] = det.get_pattern_from_pixel_coordinates_and_intensities(xy_int, inten, (70, 90), 2.5)
np.savez_compressed(HERE / "detector.npz", **out)

# ---------------------------------------------------------------- sim_utils
out = {}
for name in cases.STRUCTURES:
    st = cases.structure(name)
    for aligned in (False, True):
        sx = cases.phase(name).structure if aligned else st
        hkl = cases.hkl_box(3)
        g = sx.lattice.rnorm(hkl)
        for sp in ("lobato", "xtables", None):
            tag = f"{name}_{'orix' if aligned else 'diffpy'}_{sp}"
            out[f"I_{tag}"] = su.get_kinematical_intensities(
                sx, hkl, g, debye_waller_factors=cases.DW, scattering_params=sp, prefactor=1)
    idx, cart, dist = su.get_points_in_sphere(st.lattice.reciprocal(), 1.3)
    out[f"pts_idx_{name}"] = idx.astype(np.int32)
    out[f"pts_cart_{name}"] = cart
    out[f"pts_dist_{name}"] = dist
np.savez_compressed(HERE / "sim_utils.npz", **out)

# ---------------------------------------------------------------- old api end-to-end
out = {}
for cname, c in cases.ED_CASES.items():
    st = cases.structure(c["structure"])
    gen = dg.DiffractionGenerator(c["kv"], scattering_params=c.get("scattering_params", "lobato"),
                                  shape_factor_model=c.get("model", "lorentzian"),
                                  minimum_intensity=c.get("minimum_intensity", 1e-20))
    for i, eul in enumerate(c["eulers"]):
        sim = gen.calculate_ed_data(st, c["rr"], rotation=eul,
                                    with_direct_beam=c["with_direct_beam"],
                                    max_excitation_error=c["s_max"],
                                    debye_waller_factors=c.get("dw", {}))
        out[f"{cname}_{i}_coords"] = sim.coordinates
        out[f"{cname}_{i}_indices"] = sim.indices.astype(np.int32)
        out[f"{cname}_{i}_intensities"] = sim.intensities
        sim.calibration = c["calibration"]
        out[f"{cname}_{i}_pixel"] = np.rint(
            sim.calibrated_coordinates[:, :2] + c["half_shape"]).astype(np.int32)
        if i < 2:
            out[f"{cname}_{i}_pattern"] = sim.get_diffraction_pattern(
                shape=c["shape"], sigma=c["sigma"]).astype(np.float32)
np.savez_compressed(HERE / "ed_data.npz", **out)

# ---------------------------------------------------------------- rotation-list producers
rlg = ns.rotation_list_generators
out = {}
for system in cases.BEAM_GRID_SYSTEMS:
    g2 = rlg.get_beam_directions_grid(system, 2)
    out[f"size_{system}_2deg"] = np.array(g2.shape[0])
    if g2.shape[0] <= 2000:
        out[f"edge_{system}_2deg"] = g2
for system in ("cubic", "hexagonal", "monoclinic"):
    for mesh in ("normalized_cube", "spherified_cube_corner", "spherified_cube_edge"):
        out[f"{mesh}_{system}_5deg"] = rlg.get_beam_directions_grid(system, 5, mesh=mesh)
smg = ns.sphere_mesh_generators
for system in ("cubic", "hexagonal", "orthorhombic", "triclinic"):
    for mesh in ("uv_sphere", "icosahedral"):
        out[f"{mesh}_{system}_5deg"] = rlg.get_beam_directions_grid(system, 5, mesh=mesh)
out["vertices_uv_sphere_7deg"] = smg.get_uv_sphere_mesh_vertices(7)
out["vertices_icosahedral_9deg"] = smg.get_icosahedral_mesh_vertices(9)
out["vertices_icosahedral_3deg"] = smg.get_icosahedral_mesh_vertices(3)
out["vertices_random_4deg_seed3"] = smg.get_random_sphere_vertices(4, seed=3)
np.savez_compressed(HERE / "beam_grid.npz", **out)

# ---------------------------------------------------------------- round-2 closures (old-api DiffractionSimulation)
# (a) per-axis calibration (sims/diffraction_simulation.py:141-147); (b) knife-edge pixels: spots whose pixel coordinates are
# exact integers, where the reference's r cos(atan2(y, x) + a) + cx round trip decides the truncation by its own ulp noise
dsim = ns.diffraction_simulation
out = {}
c = cases.ED_CASES["si"]
st = cases.structure(c["structure"])
gen = dg.DiffractionGenerator(c["kv"])
sim = gen.calculate_ed_data(st, c["rr"], rotation=c["eulers"][3], with_direct_beam=True, max_excitation_error=c["s_max"])
for tag, cal, kw in (("aniso", (1.0 / 128, 1.0 / 100), dict()),
                     ("aniso_rot", (0.009, 0.0065), dict(in_plane_angle=30.0, mirrored=True))):
    sim.calibration = cal
    out[f"{tag}_coords"] = sim.coordinates
    out[f"{tag}_intensities"] = sim.intensities
    out[f"{tag}_calibration"] = np.array(cal)
    out[f"{tag}_pattern"] = sim.get_diffraction_pattern(shape=(256, 256), sigma=6, **kw).astype(np.float32)
cal = 0.01
ij = np.array([(i, j) for i in range(-4, 5) for j in range(-4, 5)], dtype=float)
coords = np.concatenate([ij * 25 * cal, np.zeros((len(ij), 1))], axis=1)          # pixels at exact multiples of 25
knife = dsim.DiffractionSimulation(coords, intensities=1.0 + np.arange(len(ij)) % 7, calibration=cal)
out["knife_coords"] = coords
out["knife_intensities"] = knife.intensities
for ang in (0.0, 90.0, 45.0):
    t = knife._get_transformed_coordinates(ang, (128, 128), False, units="pixel")
    out[f"knife_pixels_{int(ang)}"] = t[:, :2]
    out[f"knife_pattern_{int(ang)}"] = knife.get_diffraction_pattern(shape=(256, 256), sigma=2, in_plane_angle=ang).astype(np.float32)
np.savez_compressed(HERE / "closures.npz", **out)

for f in sorted(HERE.glob("*.npz")):
    print(f.name, f.stat().st_size)
