"""Pin the CPU oracle against the reference's own fixtures and against outputs of the reference's
leaf modules executed in the development container (tests/golden/make_golden.py)."""
import warnings

import numpy as np
import pytest

from oracle import kinematical as K
from tests.golden import cases


@pytest.fixture(scope="module")
def g(golden_dir):
    return {n: np.load(golden_dir / f"{n}.npz") for n in
            ("old_simulation", "shape_factors", "detector", "sim_utils", "ed_data")}


def test_wavelengths():
    # diffsims/tests/utils/test_sim_utils.py:49-60
    for kv, lam in [(100, 0.0370143659), (200, 0.0250793403), (300, 0.0196874888)]:
        np.testing.assert_almost_equal(K.get_electron_wavelength(kv), lam)
    assert K.get_electron_wavelength(np.inf) == 0


def test_graphite_golden_new_and_old_api(g):
    """diffsims/tests/generators/test_simulation_generator.py:283-343 (atol 1e-8)."""
    old = g["old_simulation"]["image"]
    p = cases.phase("graphite")
    gs = K.GSet(p.structure, 1.6768, False)
    assert len(gs.xyz) == 664
    wl = K.get_electron_wavelength(200)
    r = K.simulate_rotation(p.structure, gs, K.bunge_matrix(*np.deg2rad([0, 90, 90])), wl, 0.1)
    assert len(r["intensity"]) == 104
    img = K.diffraction_pattern(r["xyz"], r["intensity"], (128, 128), sigma=1.4, calibration=0.0262)
    np.testing.assert_allclose(img, old, atol=1e-8)
    st = cases.structure("graphite")
    r2 = K.calculate_ed_data(st, 1.6768, wl, rotation=(0, 90, 120), with_direct_beam=False,
                             max_excitation_error=0.1)
    m = r2["direct_beam_mask"]
    img2 = K.old_diffraction_pattern(r2["coordinates"][m], r2["intensities"][m], 0.0262, (128, 128), 1.4)
    np.testing.assert_allclose(img2, old, atol=1e-8)


def test_reference_reflection_counts():
    """test_simulation_generator.py:137-168: 70, 250 (both precession modes), 52 (custom model)."""
    si = cases.phase("si")
    wl = K.get_electron_wavelength(300)
    gs = K.GSet(si.structure, 5.0, True)
    I3 = np.eye(3)
    assert len(K.simulate_rotation(si.structure, gs, I3, wl)["intensity"]) == 70
    assert len(K.simulate_rotation(si.structure, gs, I3, wl, precession_angle=0.5)["intensity"]) == 250

    def local_excite(s, m, t):
        return (np.sin(t) * s) / m

    gsr = K.GSet(si.structure, 5.0, True, emulate_orix_rounding=True)
    r = K.simulate_rotation(si.structure, gsr, I3, wl, shape_factor_model=local_excite,
                            shape_factor_kwargs=dict(t=0.2))
    assert len(r["intensity"]) == 52   # 36 allowed + 16 round-off "forbidden" reflections
    r = K.simulate_rotation(si.structure, gs, I3, wl, shape_factor_model=local_excite,
                            shape_factor_kwargs=dict(t=0.2))
    assert len(r["intensity"]) == 36   # with exact integer hkl (what the device path uses)


@pytest.mark.slow
def test_reference_count_full_precession():
    si = cases.phase("si")
    gs = K.GSet(si.structure, 5.0, True)
    r = K.simulate_rotation(si.structure, gs, np.eye(3), K.get_electron_wavelength(300),
                            precession_angle=0.5, approximate_precession=False)
    assert len(r["intensity"]) == 250


def test_docstring_structure_factors():
    """reciprocal_lattice_vector.py:542, :680, :690 (Al a = 4.04) and test_sim_utils.py:330-353 (Ni)."""
    from diffsims_b200.crystal import Atom, Lattice, Phase, Structure
    al = Phase("al", space_group=225, structure=Structure(
        [Atom("Al", p) for p in ([0, 0, 1], [.5, .5, 1], [.5, 0, .5], [0, .5, .5])],
        Lattice(4.04, 4.04, 4.04, 90, 90, 90)))
    hkl = np.array([[1, 1, 1], [2, 0, 0]])
    gn = al.structure.lattice.rnorm(hkl)
    np.testing.assert_allclose(np.abs(K.kinematical_structure_factor(al.structure, hkl, gn, None, "xtables")),
                               [8.46881663, 7.04777513], rtol=1e-8)
    np.testing.assert_allclose(np.abs(K.kinematical_structure_factor(al.structure, hkl, gn, None, "lobato")),
                               [8.44934816, 7.0387957], rtol=1e-8)
    ni = Structure([Atom("Ni", [0, 0, 1])], Lattice(3.5, 3.5, 3.5, 90, 90, 90))
    np.testing.assert_allclose(K.kinematical_intensities(ni, np.array([[0, 0, 0]]), np.array([0.0])),
                               [43.0979], rtol=1e-5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_allclose(
            K.kinematical_intensities(ni, np.array([[0, 0, 0]]), np.array([0.0]), scattering_params=None), [1.0])


def test_g_set_sizes_and_order():
    """docstring reciprocal_lattice_vector.py:1109, :1124-1125; tests/crystallography :89-116."""
    from diffsims_b200.crystal import Atom, Lattice, Phase, Structure
    al = Phase("al", space_group=225, structure=Structure([Atom("Al", [0, 0, 1])],
                                                          Lattice(4.04, 4.04, 4.04, 90, 90, 90)))
    hkl, _, _ = K.from_min_dspacing(al.structure.lattice, 0.7)
    assert len(hkl) == 798 and tuple(hkl[0]) == (5, 2, 2) and tuple(hkl[-1]) == (-5, -2, -2)
    assert len(K.from_min_dspacing(al.structure.lattice, 1.0)[0]) == 256
    hkl0 = K.from_min_dspacing(al.structure.lattice, 1.0, include_zero_vector=True)[0]
    assert len(hkl0) == 257 and tuple(hkl0[-1]) == (0, 0, 0)


def test_shape_factors_against_reference(g):
    sf = g["shape_factors"]
    s = sf["s"]
    for w in (0.01, 0.05, 0.1):
        np.testing.assert_allclose(K.linear(s.copy(), w), sf[f"linear_{w}"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(K.sinc(s.copy(), w), sf[f"sinc_{w}"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(K.sinc(s.copy(), w, 7), sf[f"sinc7_{w}"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(K.sin2c(s.copy(), w), sf[f"sin2c_{w}"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(K.atanc(s.copy(), w), sf[f"atanc_{w}"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(K.lorentzian(s.copy(), w), sf[f"lorentzian_{w}"], rtol=1e-13)
        np.testing.assert_allclose(K.lorentzian_precession(s.copy(), w, sf["r_spot"], np.deg2rad(0.5)),
                                   sf[f"lorentzian_precession_{w}"], rtol=1e-13)
    # the reference's quirks at s == 0
    assert K.sinc(np.array([0.0]), 0.01)[0] == 0 and K.atanc(np.array([0.0]), 0.01)[0] == 1


@pytest.mark.parametrize("name", list(cases.DETECTOR_CASES))
def test_rasteriser_against_reference(g, name):
    shape, sigma, n, seed = cases.DETECTOR_CASES[name]
    xy, inten = cases.detector_spots(shape, n, seed)
    np.testing.assert_allclose(
        K.pattern_from_pixel_coordinates_and_intensities(xy.astype(int), inten, shape, sigma),
        g["detector"][f"{name}_int"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(
        K.pattern_from_pixel_coordinates_and_intensities(xy, inten * 2000.0, shape, sigma, 1.0),
        g["detector"][f"{name}_float"], rtol=1e-12, atol=1e-15)


def test_rasteriser_spots_outside_the_frame_against_reference(g):
    """The bare function spreads out-of-frame float spots into the frame (incl. numpy's wrap of a negative slice
    stop) and wraps negative integer indices."""
    xy, xy_int, inten = cases.detector_spots_outside((70, 90), 40, 5)
    np.testing.assert_allclose(K.pattern_from_pixel_coordinates_and_intensities(xy, inten, (70, 90), 2.5),
                               g["detector"]["outside_float"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(K.pattern_from_pixel_coordinates_and_intensities(xy_int, inten, (70, 90), 2.5),
                               g["detector"]["outside_int"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name", list(cases.STRUCTURES))
def test_intensities_against_reference(g, name):
    hkl = cases.hkl_box(3)
    for aligned in (False, True):
        sx = cases.phase(name).structure if aligned else cases.structure(name)
        gn = sx.lattice.rnorm(hkl)
        for sp in ("lobato", "xtables", None):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = K.kinematical_intensities(sx, hkl, gn, cases.DW, sp)
            ref = g["sim_utils"][f"I_{name}_{'orix' if aligned else 'diffpy'}_{sp}"]
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * ref.max())
    st = cases.structure(name)
    idx, cart, dist = K.points_in_sphere(st.lattice.reciprocal(), 1.3)
    np.testing.assert_array_equal(idx.astype(np.int32), g["sim_utils"][f"pts_idx_{name}"])
    np.testing.assert_allclose(cart, g["sim_utils"][f"pts_cart_{name}"], atol=1e-14)
    np.testing.assert_allclose(dist, g["sim_utils"][f"pts_dist_{name}"], atol=1e-14)


@pytest.mark.parametrize("cname", list(cases.ED_CASES))
def test_old_api_against_reference(g, cname):
    """ed_data.npz: the reference's DiffractionGenerator.calculate_ed_data executed on these inputs."""
    c = cases.ED_CASES[cname]
    st = cases.structure(c["structure"])
    wl = K.get_electron_wavelength(c["kv"])
    e = g["ed_data"]
    for i, eul in enumerate(c["eulers"]):
        r = K.calculate_ed_data(st, c["rr"], wl, rotation=eul, with_direct_beam=c["with_direct_beam"],
                                max_excitation_error=c["s_max"],
                                shape_factor_model=K.SHAPE_FACTOR_MODELS[c.get("model", "lorentzian")],
                                scattering_params=c.get("scattering_params", "lobato"),
                                debye_waller_factors=c.get("dw", {}),
                                minimum_intensity=c.get("minimum_intensity", 1e-20))
        m = r["direct_beam_mask"]
        np.testing.assert_array_equal(r["indices"][m].astype(np.int32), e[f"{cname}_{i}_indices"])
        np.testing.assert_allclose(r["coordinates"][m], e[f"{cname}_{i}_coords"], atol=1e-13)
        np.testing.assert_allclose(r["intensities"][m], e[f"{cname}_{i}_intensities"], rtol=1e-10)
        np.testing.assert_array_equal(
            K.library_pixel_coords(r["coordinates"][m], c["calibration"], c["half_shape"]),
            e[f"{cname}_{i}_pixel"])
        if i < 2:
            img = K.old_diffraction_pattern(r["coordinates"][m], r["intensities"][m], c["calibration"],
                                            c["shape"], c["sigma"])
            np.testing.assert_allclose(img, e[f"{cname}_{i}_pattern"], atol=2e-7)


def test_beam_directions_grid_against_reference(golden_dir):
    """Rotation-list producer: sizes pinned by the reference's tests
    (tests/generators/test_rotation_list_generator.py:80-94) and full arrays from the reference itself."""
    gold = np.load(golden_dir / "beam_grid.npz")
    for system in cases.BEAM_GRID_SYSTEMS:
        got = K.beam_directions_grid(system, 2)
        assert got.shape[0] == cases.BEAM_GRID_SIZES_2DEG[system] == int(gold[f"size_{system}_2deg"])
        if f"edge_{system}_2deg" in gold.files:
            np.testing.assert_array_equal(got, gold[f"edge_{system}_2deg"])
    for key in gold.files:
        if key.endswith("_5deg"):
            mesh, system = key[:-5].rsplit("_", 1)
            np.testing.assert_array_equal(K.beam_directions_grid(system, 5, mesh=mesh), gold[key])
    # mesh vertices of sphere_mesh_generators.py (uv sphere :42-93, icosahedral :378-450, random :453-483)
    np.testing.assert_array_equal(K.uv_sphere_mesh_vertices(7), gold["vertices_uv_sphere_7deg"])
    np.testing.assert_array_equal(K.icosahedral_mesh_vertices(9), gold["vertices_icosahedral_9deg"])
    np.testing.assert_array_equal(K.icosahedral_mesh_vertices(3), gold["vertices_icosahedral_3deg"])
    np.testing.assert_array_equal(K.random_sphere_vertices(4, seed=3), gold["vertices_random_4deg_seed3"])
    # diffsims/tests/generators/test_sphere_mesh_generators.py: shapes only
    assert K.uv_sphere_mesh_vertices(10).shape[1] == 3 and K.icosahedral_mesh_vertices(10).shape[1] == 3
