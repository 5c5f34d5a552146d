"""The reference's own hot-path tests, restated against the drop-in API (GPU), plus parity of the API
results with the CPU oracle and with golden outputs of the reference's leaf modules."""
import re
import warnings

import numpy as np
import pytest

import diffsims_b200 as ds
from diffsims_b200.crystal import Atom, Lattice, Phase, Rotation, Structure
from diffsims_b200.crystallography import DiffractingVector
from diffsims_b200.simulations import Simulation2D
from diffsims_b200.utils import shape_factor_models as sfm
from diffsims_b200.utils.sim_utils import get_kinematical_intensities, get_kinematical_structure_factor
from oracle import kinematical as K
from tests.golden import cases
from tests.helpers import IMG_ATOL, RTOL, compare_spots

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def make_phase(a=5.431):
    latt = Lattice(a, a, a, 90, 90, 90)
    atoms = []
    for c in [[0, 0, 0], [0.5, 0, 0.5], [0, 0.5, 0.5], [0.5, 0.5, 0]]:
        atoms.append(Atom(atype="Si", xyz=c, lattice=latt))
        atoms.append(Atom(atype="Si", xyz=[c[0] + 0.25, c[1] + 0.25, c[2] + 0.25], lattice=latt))
    return Phase(structure=Structure(atoms=atoms, lattice=latt), space_group=227)


# ---------------------------------------------------------------- test_simulation_generator.py restated
class TestDiffractionCalculator:
    def test_matching_results(self):   # :137-143
        d = ds.SimulationGenerator(300).calculate_diffraction2d(make_phase(), reciprocal_radius=5.0)
        assert isinstance(d.coordinates, DiffractingVector)
        assert d.coordinates.size == 70

    def test_precession_simple(self):  # :145-152
        gen = ds.SimulationGenerator(300, precession_angle=0.5, approximate_precession=True)
        assert gen.calculate_diffraction2d(make_phase(), reciprocal_radius=5.0).coordinates.size == 250

    def test_precession_full(self):  # :154-161
        gen = ds.SimulationGenerator(300, precession_angle=0.5, approximate_precession=False)
        assert gen.calculate_diffraction2d(make_phase(), reciprocal_radius=5.0).coordinates.size == 250

    @pytest.mark.parametrize("model", ["lorentzian", "linear", "sinc", "sin2c", "atanc", sfm.binary])
    def test_precession_full_matches_oracle_quad(self, model):
        """K2's midpoint-rule average over the precession circle vs the reference's scipy.quad integral."""
        phase = cases.phase("si")
        gen = ds.SimulationGenerator(300, precession_angle=0.5, approximate_precession=False,
                                     shape_factor_model=model)
        rot = Rotation.from_euler([[0, 0, 0], [12, 34, 56]], degrees=True)
        sim = gen.calculate_diffraction2d(phase, rot, reciprocal_radius=1.6, max_excitation_error=0.02)
        gs = K.GSet(phase.structure, 1.6, True)
        name = model if isinstance(model, str) else "binary"
        # scipy.quad (the reference) hits its 50-subdivision limit on the kinked models and is then good to
        # ~1e-4 only (its own error estimate); a converged midpoint rule is the tight check for those
        quad_rtol = RTOL if name in ("lorentzian", "atanc", "binary") else 2e-3
        for i, dv in enumerate(sim):
            got = {tuple(h.astype(int)): v for h, v in zip(dv.hkl, dv.intensity)}
            for converged, rtol in ((False, quad_rtol), (True, 2e-6)):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref = K.simulate_rotation(phase.structure, gs, rot.to_matrix()[i], gen.wavelength, 0.02,
                                              shape_factor_model=K.SHAPE_FACTOR_MODELS[name], precession_angle=0.5,
                                              approximate_precession=False, converged_precession=converged)
                big = ref["intensity"].max()
                keep = ref["intensity"] > 1e-10 * big
                assert dv.size >= keep.sum()
                for h, v in zip(ref["hkl_int"][keep], ref["intensity"][keep]):
                    assert abs(got[tuple(h)] - v) <= rtol * v + 1e-9 * big, (converged, h, got[tuple(h)], v)

    def test_precession_full_needs_native_model(self):
        gen = ds.SimulationGenerator(300, precession_angle=0.5, approximate_precession=False,
                                     shape_factor_model=lambda s, w: s)
        with pytest.raises(NotImplementedError):
            gen.calculate_diffraction2d(make_phase(), reciprocal_radius=1.0)

    def test_custom_shape_func(self):  # :163-168
        def local_excite(excitation_error, maximum_excitation_error, t):
            return (np.sin(t) * excitation_error) / maximum_excitation_error

        gen = ds.SimulationGenerator(300, shape_factor_model=local_excite, t=0.2)
        d = gen.calculate_diffraction2d(make_phase(), reciprocal_radius=5.0)
        # the reference pins 52 = 36 allowed reflections with s > 0 + 16 forbidden reflections that survive
        # only as 1e-10 rounding noise of orix' unique(); with exact integer hkl the 36 remain
        # (oracle: tests/test_oracle_golden.py::test_reference_reflection_counts)
        assert d.coordinates.size == 36
        assert np.all(d.coordinates.intensity > 0)

    def test_appropriate_scaling(self):  # :170-186
        gen = ds.SimulationGenerator(300)
        d = gen.calculate_diffraction2d(phase=make_phase(5), reciprocal_radius=5.0)
        big = gen.calculate_diffraction2d(phase=make_phase(10), reciprocal_radius=5.0)
        idx = [tuple(i) for i in d.coordinates.hkl]
        big_idx = [tuple(i) for i in big.coordinates.hkl]
        assert (2, 2, 0) in idx and (2, 2, 0) in big_idx
        c = d.coordinates[idx.index((2, 2, 0))]
        bc = big.coordinates[big_idx.index((2, 2, 0))]
        assert np.allclose(c.data, bc.data * 2)

    def test_appropriate_intensities(self):  # :188-200
        d = ds.SimulationGenerator(300).calculate_diffraction2d(make_phase(), reciprocal_radius=0.5,
                                                                with_direct_beam=True)
        idx = [tuple(np.round(i).astype(int)) for i in d.coordinates.hkl]
        central = idx.index((0, 0, 0))
        assert np.all(np.greater_equal(d.coordinates.intensity[central], d.coordinates.intensity))

    def test_direct_beam(self):  # :202-208
        d = ds.SimulationGenerator(300).calculate_diffraction2d(make_phase(), reciprocal_radius=0.5,
                                                                with_direct_beam=False)
        idx = [tuple(np.round(i).astype(int)) for i in d.coordinates.hkl]
        assert (0, 0, 0) not in idx

    def test_shape_factor_custom(self):  # :215-224
        gen = ds.SimulationGenerator(300)
        t1 = gen.calculate_diffraction2d(make_phase(), max_excitation_error=0.02)
        t2 = gen.calculate_diffraction2d(make_phase(), max_excitation_error=0.4)
        assert np.sum(t1.coordinates.intensity) != np.sum(t2.coordinates.intensity)

    @pytest.mark.parametrize("model", ["linear", "atanc", "sinc", "sin2c", "lorentzian", sfm.binary])
    def test_every_native_model_runs(self, model):
        gen = ds.SimulationGenerator(300, shape_factor_model=model)
        assert gen.calculate_diffraction2d(make_phase()).coordinates.size > 0


def test_multiphase_multirotation_simulation():  # :249-256
    gen = ds.SimulationGenerator(300)
    rot = Rotation.from_euler([[0, 0, 0], [0.1, 0.1, 0.1]])
    rot2 = Rotation.from_euler([[0, 0, 0], [0.1, 0.1, 0.1], [0.2, 0.2, 0.2]])
    sim = gen.calculate_diffraction2d([make_phase(5), make_phase(10)], rotation=[rot, rot2])
    assert sim.num_phases == 2 and [len(c) for c in sim.coordinates] == [2, 3]
    assert sum(1 for _ in sim) == 5
    assert isinstance(sim.iphase[1].irot[2].coordinates, DiffractingVector)


def test_multiphase_multirotation_simulation_error():  # :258-265
    gen = ds.SimulationGenerator(300)
    rot = Rotation.from_euler([[0, 0, 0], [0.1, 0.1, 0.1]])
    with pytest.raises(ValueError):
        gen.calculate_diffraction2d([make_phase(5), make_phase(10)], rotation=[rot])


def test_same_simulation_results(golden_dir):  # :283-343, the reference's golden image
    latt = Lattice(2.464, 2.464, 6.711, 90, 90, 120)
    atoms = [Atom(atype="C", xyz=[0.0, 0.0, 0.25], lattice=latt), Atom(atype="C", xyz=[0.0, 0.0, 0.75], lattice=latt),
             Atom(atype="C", xyz=[1 / 3, 2 / 3, 0.25], lattice=latt), Atom(atype="C", xyz=[2 / 3, 1 / 3, 0.75], lattice=latt)]
    structure_matrix = Structure(atoms=atoms, lattice=latt)
    kw = dict(accelerating_voltage=200, scattering_params="lobato", precession_angle=0,
              shape_factor_model="lorentzian", approximate_precession=True, minimum_intensity=1e-20)
    p = Phase("Graphite", point_group="6/mmm", structure=structure_matrix)
    sim = ds.SimulationGenerator(**kw).calculate_diffraction2d(
        phase=p, rotation=Rotation.from_euler(np.array([[0, 90, 90]]), degrees=True),
        reciprocal_radius=1.6768, max_excitation_error=0.1, with_direct_beam=False)
    new_data = sim.get_diffraction_pattern(shape=(128, 128), sigma=1.4, calibration=0.0262)
    old_data = np.load(golden_dir / "old_simulation.npz")["image"]
    assert new_data.dtype == np.float64 and new_data.shape == (128, 128)
    np.testing.assert_allclose(new_data, old_data, atol=IMG_ATOL)  # float32 raster: 1e-4 of peak

    # and the same pattern through the OLD api (the commented-out generator of the golden, :314-325)
    lib = ds.DiffractionLibraryGenerator(ds.DiffractionGenerator(**kw)).get_diffraction_library(
        ds.StructureLibrary(["Graphite"], [structure_matrix], [np.array([[0, 90, 120]])]),
        calibration=0.0262, reciprocal_radius=1.6768, with_direct_beam=False, max_excitation_error=0.1,
        half_shape=64)
    old_way = lib["Graphite"]["simulations"][0].get_diffraction_pattern(shape=(128, 128), sigma=1.4)
    np.testing.assert_allclose(old_way, old_data, atol=IMG_ATOL)


def test_calculate_diffraction2d_progressbar(capsys):  # :346-395
    gen = ds.SimulationGenerator()
    phase = make_phase()
    phase.name = "test phase"
    rots = Rotation.random(10)
    gen.calculate_diffraction2d(phase, rots, show_progressbar=False)
    assert capsys.readouterr().err == ""
    gen.calculate_diffraction2d(phase, rots, show_progressbar=True)
    assert re.findall(r"test phase: 100\%\|█+\| 10\/10", capsys.readouterr().err)
    p2 = make_phase()
    p2.name = "B"
    gen.calculate_diffraction2d([phase, p2], [rots, rots], show_progressbar=True)
    err = capsys.readouterr().err
    assert re.findall(r"test phase: 100\%\|█+\| 10\/10 ", err) and re.findall(r"B: 100\%\|█+\| 10\/10 ", err)


# ---------------------------------------------------------------- API results == oracle
@pytest.mark.parametrize("name,kv,rr,s_max,db", [("si", 200, 1.0, 0.01, True), ("ti", 300, 1.2, 0.02, False),
                                               ("fe3c", 200, 1.0, 0.02, True), ("triclinic", 120, 0.9, 0.02, True)])
def test_calculate_diffraction2d_matches_oracle(name, kv, rr, s_max, db):
    phase = cases.phase(name)
    rot = Rotation.random(16, rng=3)
    gen = ds.SimulationGenerator(kv)
    sim = gen.calculate_diffraction2d(phase, rot, reciprocal_radius=rr, max_excitation_error=s_max,
                                      with_direct_beam=db, debye_waller_factors=cases.DW)
    gs = K.GSet(phase.structure, rr, db)
    G = rot.to_matrix()
    hkl_ref = np.vstack([gs.hkl_int, [0, 0, 0]]) if db else gs.hkl_int
    for i, dv in enumerate(sim):
        ref = K.simulate_rotation(phase.structure, gs, G[i], K.get_electron_wavelength(kv), s_max,
                                  debye_waller_factors=cases.DW)
        # key the API's reflections by their position in the oracle's table via hkl
        key = {tuple(h): j for j, h in enumerate(hkl_ref)}
        gidx = [key[tuple(h.astype(int))] for h in dv.hkl]
        if db:  # the duplicated (000): second occurrence is the extra row
            z = [j for j, h in enumerate(dv.hkl) if not h.any()]
            if len(z) == 2:
                gidx[z[0]], gidx[z[1]] = len(hkl_ref) - 2, len(hkl_ref) - 1
        got = dict(g_index=np.array(gidx), xyz=dv.data, intensity=dv.intensity,
                   excitation_error=np.zeros(dv.size))
        compare_spots(ref, got, s_max=-1.0, rr=rr, prec=True)  # strict set equality (no cut-edge cases here)
        # rotated basis: hkl recomputed from the rotated lattice agrees with the table's integers
        np.testing.assert_allclose(dv.data @ dv.phase.structure.lattice.base.T, dv.hkl, atol=1e-9)


def test_get_intersecting_reflections_matches_oracle():
    phase = cases.phase("si")
    gen = ds.SimulationGenerator(200)
    recip = DiffractingVector.from_min_dspacing(phase, min_dspacing=1.0, include_zero_vector=True)
    rot = Rotation.from_euler([[10, 20, 30]], degrees=True)
    dv, hkl, sf = gen.get_intersecting_reflections(recip, rot, gen.wavelength, 0.02)
    gs = K.GSet(phase.structure, 1.0, True)
    rotated = np.vstack([gs.xyz @ rot.to_matrix()[0], [0, 0, 0]])
    xyz, hkl_ref, sf_ref, s, idx = K.intersecting_reflections(
        rotated, np.vstack([gs.hkl_float, [0, 0, 0]]), gen.wavelength, 0.02)
    np.testing.assert_allclose(dv.data, xyz, atol=1e-12)
    np.testing.assert_allclose(hkl, hkl_ref, atol=1e-9)
    np.testing.assert_allclose(sf, sf_ref, rtol=1e-9)


@pytest.mark.parametrize("name", ["si", "fe3c", "triclinic"])
def test_get_kinematical_intensities_matches_reference_golden(golden_dir, name):
    g = np.load(golden_dir / "sim_utils.npz")
    hkl = cases.hkl_box(3)
    for aligned in (False, True):
        sx = cases.phase(name).structure if aligned else cases.structure(name)
        gn = sx.lattice.rnorm(hkl)
        for sp in ("lobato", "xtables", None):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = get_kinematical_intensities(sx, hkl, gn, debye_waller_factors=cases.DW, scattering_params=sp)
            ref = g[f"I_{name}_{'orix' if aligned else 'diffpy'}_{sp}"]
            np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12 * ref.max())


def test_get_kinematical_intensities_reference_tests():
    """diffsims/tests/utils/test_sim_utils.py:330-394."""
    ni = Structure([Atom("Ni", [0, 0, 1])], Lattice(3.5, 3.5, 3.5, 90, 90, 90))
    hkl, gn = np.array([[0, 0, 0]]), np.array([0.0])
    np.testing.assert_array_almost_equal(
        get_kinematical_intensities(ni, hkl, gn, prefactor=1, scattering_params="lobato"), [43.0979], decimal=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_array_almost_equal(
            get_kinematical_intensities(ni, hkl, gn, prefactor=1, scattering_params=None), [1.0])
    with pytest.raises(NotImplementedError):
        get_kinematical_intensities(ni, hkl, gn, scattering_params="_empty")
    bad = Structure([Atom("Zz", [0, 0, 0])], Lattice(3.5, 3.5, 3.5, 90, 90, 90))
    with pytest.warns(UserWarning, match="not found in scattering parameter library"):
        np.testing.assert_allclose(get_kinematical_intensities(bad, hkl, gn), [0.0])
    al = Phase("al", space_group=225, structure=Structure(
        [Atom("Al", p) for p in ([0, 0, 1], [.5, .5, 1], [.5, 0, .5], [0, .5, .5])], Lattice(4.04, 4.04, 4.04, 90, 90, 90)))
    h2 = np.array([[1, 1, 1], [2, 0, 0]])
    F = get_kinematical_structure_factor(al.structure, h2, al.structure.lattice.rnorm(h2), scattering_params="xtables")
    np.testing.assert_allclose(np.abs(F), [8.46881663, 7.04777513], rtol=1e-8)
    np.testing.assert_allclose(get_kinematical_intensities(ni, hkl, gn, prefactor=np.array([2.0])), [2 * 43.09791], rtol=1e-6)


# ---------------------------------------------------------------- rendering through the API
def test_get_diffraction_pattern_fast_vs_slow():
    """diffsims/tests/simulations/test_simulations2d.py:133-180."""
    al = Phase(name="al", space_group=225,
               structure=Structure(atoms=[Atom("al", [0, 0, 0])], lattice=Lattice(0.405, 0.405, 0.405, 90, 90, 90)))
    gen = ds.SimulationGenerator()
    rot = Rotation.identity()
    xyz, inten = np.asarray([[0, 0, 0]]), np.array([30])
    int_sim = Simulation2D(phases=al, simulation_generator=gen, rotations=rot,
                           coordinates=DiffractingVector(phase=al, xyz=xyz.astype(int), intensity=inten))
    float_sim = Simulation2D(phases=al, simulation_generator=gen, rotations=rot,
                             coordinates=DiffractingVector(phase=al, xyz=xyz.astype(float), intensity=inten))
    kw = dict(shape=(11, 11), sigma=1, calibration=1, normalize=False, clip_threshold=0.01)
    fast = int_sim.get_diffraction_pattern(fast=True, **kw)
    slow = int_sim.get_diffraction_pattern(fast=False, **kw)
    assert np.array_equal(fast, slow)          # integer coordinates take the integer branch either way
    float_fast = float_sim.get_diffraction_pattern(fast=True, **kw)
    float_slow = float_sim.get_diffraction_pattern(fast=False, **kw)
    assert np.allclose(float_fast, fast)
    assert np.all((float_slow - float_fast) < kw["clip_threshold"])
    kw["shape"] = (10, 10)                     # centre between pixels: only the slow path sees it
    assert not np.allclose(float_sim.get_diffraction_pattern(fast=True, **kw),
                           float_sim.get_diffraction_pattern(fast=False, **kw))
    # against the oracle
    ref = K.diffraction_pattern(xyz.astype(float), inten.astype(float), (10, 10), sigma=1, calibration=1,
                                fast=False, normalize=False, clip_threshold=0.01)
    np.testing.assert_allclose(float_sim.get_diffraction_pattern(fast=False, **kw), ref, atol=IMG_ATOL * ref.max())


def test_get_diffraction_pattern_empty_and_normalised():
    """test_simulations2d.py:457-470."""
    al = Phase(name="al", space_group=225,
               structure=Structure(atoms=[Atom("al", [0, 0, 0])], lattice=Lattice(0.405, 0.405, 0.405, 90, 90, 90)))
    gen = ds.SimulationGenerator(accelerating_voltage=200)
    rot = Rotation.from_euler([[0, a, 0] for a in (0, 15, 30, 45)], degrees=True)
    coords = DiffractingVector(phase=al, xyz=[[1, 0, 0], [0, -0.3, 0], [1 / 0.405, 1 / -0.405, 0], [0.1, -0.1, -0.3]])
    coords.intensity = 1
    p2 = al.deepcopy()
    p2.name = "al2"
    sim = Simulation2D(phases=[al, p2], simulation_generator=gen, coordinates=[[coords] * 4, [coords] * 4],
                       rotations=[rot, rot])
    pat = sim.get_diffraction_pattern(shape=(50, 50), calibration=0.001)
    assert pat.shape == (50, 50) and np.max(pat) == 0
    pat = sim.get_diffraction_pattern(shape=(512, 512), calibration=0.01)
    assert pat.shape == (512, 512) and np.max(pat) == 1


def test_batched_patterns_equal_single_and_oracle():
    phase = cases.phase("si")
    rot = Rotation.random(12, rng=5)
    sim = ds.SimulationGenerator(200).calculate_diffraction2d(phase, rot, reciprocal_radius=1.0,
                                                              max_excitation_error=0.02)
    kw = dict(shape=(256, 256), sigma=10, calibration=1 / 128)
    batch = sim.get_diffraction_patterns(**kw).cpu().numpy()
    assert batch.shape == (12, 256, 256) and batch.dtype == np.float32
    for i in (0, 5, 11):
        sim.rotation_index = i
        single = sim.get_diffraction_pattern(**kw)
        np.testing.assert_allclose(single, batch[i], atol=3e-5)   # (possibly another K3 kernel, see below)
        dv = sim.coordinates[i]
        ref = K.diffraction_pattern(dv.data, dv.intensity, **kw)
        assert np.abs(batch[i] - ref).max() <= IMG_ATOL
    sub = sim.irot[3:7].get_diffraction_patterns(**kw).cpu().numpy()
    # (a sub-list has its own row capacity and may take another K3 kernel: float32 gather vs tcgen05 split products,
    # which agree to ~2e-5 of the peak, not bit for bit)
    assert np.abs(sub - batch[3:7]).max() <= 3e-5
    sim.rotation_index = 0   # iteration is stateful, as in the reference
    r, t, inten = sim.polar_flatten_simulations()
    assert r.shape[0] == 12 and r.shape == t.shape == inten.shape


# ---------------------------------------------------------------- old api against reference-executed goldens
@pytest.mark.parametrize("cname", list(cases.ED_CASES))
def test_old_api_matches_reference_golden(golden_dir, cname):
    c = cases.ED_CASES[cname]
    e = np.load(golden_dir / "ed_data.npz")
    st = cases.structure(c["structure"])
    gen = ds.DiffractionGenerator(c["kv"], scattering_params=c.get("scattering_params", "lobato"),
                                  shape_factor_model=c.get("model", "lorentzian"),
                                  minimum_intensity=c.get("minimum_intensity", 1e-20))
    lib = ds.DiffractionLibraryGenerator(gen).get_diffraction_library(
        ds.StructureLibrary([cname], [st], [c["eulers"]]), calibration=c["calibration"],
        reciprocal_radius=c["rr"], half_shape=c["half_shape"], with_direct_beam=c["with_direct_beam"],
        max_excitation_error=c["s_max"], debye_waller_factors=c.get("dw", {}))
    assert isinstance(lib, ds.DiffractionLibrary) and lib.reciprocal_radius == c["rr"]
    entry = lib[cname]
    for i, eul in enumerate(c["eulers"]):
        sim = entry["simulations"][i]
        ref_idx = e[f"{cname}_{i}_indices"]
        ref_I = e[f"{cname}_{i}_intensities"]
        big = ref_I.max()
        # identical reflection sets except round-off "reflections" (none are near the cut in these cases)
        rk = {tuple(h): j for j, h in enumerate(ref_idx)}
        gk = {tuple(h): j for j, h in enumerate(sim.indices.astype(int))}
        for k in set(rk) ^ set(gk):
            I = ref_I[rk[k]] if k in rk else sim.intensities[gk[k]]
            assert I < 1e-10 * big, (k, I)
        common = [k for k in rk if k in gk]
        ri, gi = [rk[k] for k in common], [gk[k] for k in common]
        keep = ref_I[ri] >= 1e-10 * big
        np.testing.assert_allclose(sim.intensities[gi][keep], ref_I[ri][keep], rtol=RTOL)
        np.testing.assert_allclose(sim.coordinates[gi], e[f"{cname}_{i}_coords"][ri], rtol=RTOL, atol=1e-6 * c["rr"])
        # single-orientation entry point agrees with the batched one
        one = gen.calculate_ed_data(st, c["rr"], rotation=eul, with_direct_beam=c["with_direct_beam"],
                                    max_excitation_error=c["s_max"], debye_waller_factors=c.get("dw", {}))
        np.testing.assert_array_equal(one.indices, sim.indices)
        np.testing.assert_array_equal(one.intensities, sim.intensities)
        if len(common) == len(rk) == len(gk):
            px = e[f"{cname}_{i}_pixel"]
            # rint() can flip at exact half pixels; everywhere else identical
            xy = sim.calibrated_coordinates[:, :2] + c["half_shape"]
            safe = (np.abs(xy - np.floor(xy) - 0.5) > 1e-6).all(axis=1)
            np.testing.assert_array_equal(entry["pixel_coords"][i][gi][safe[gi]], px[ri][safe[gi]])
        if i < 2:
            img = sim.get_diffraction_pattern(shape=c["shape"], sigma=c["sigma"])
            assert np.abs(img - e[f"{cname}_{i}_pattern"]).max() <= IMG_ATOL
    le = lib.get_library_entry(phase=cname, angle=c["eulers"][1])
    assert le["Sim"] is entry["simulations"][1]


# ---------------------------------------------------------------- polar flattening on the packed result
@pytest.mark.parametrize("axes", [False, True])
def test_polar_flatten_packed_matches_object_path(axes):
    """The device kernel (ds_polar_flatten) against the reference's per-template loop
    (simulation2d.py:313-355), which still runs when the result is a plain list of DiffractingVectors."""
    gen = ds.SimulationGenerator(200)
    rots = [Rotation.random(9, rng=1), Rotation.random(5, rng=2)]
    phases = [cases.phase("si"), cases.phase("ti")]
    kw = {}
    if axes:
        kw = dict(radial_axes=np.linspace(0, 1.2, 40), azimuthal_axes=np.linspace(-np.pi, np.pi, 90))
    for sim in (gen.calculate_diffraction2d(phases[0], rots[0], max_excitation_error=0.03),
                gen.calculate_diffraction2d(phases, rots, max_excitation_error=0.03)):
        fast = sim.polar_flatten_simulations(**kw)
        # same data as plain objects -> the reference's loop
        if sim.has_multiple_phases:
            coords = [[c for c in pv] for pv in sim.coordinates]
        else:
            coords = [c for c in sim.coordinates]
        plain = Simulation2D(phases=sim.phases, coordinates=coords, rotations=sim.rotations,
                             simulation_generator=gen)
        assert plain._packed_phases() is None
        slow = plain.polar_flatten_simulations(**kw)
        if not axes:   # independent numpy evaluation of to_flat_polar (_diffracting_vector.py:186-194)
            vecs = [v for pv in sim.coordinates for v in pv] if sim.has_multiple_phases else list(sim.coordinates)
            for row, v in enumerate(vecs):
                np.testing.assert_allclose(fast[0][row, : v.size], np.hypot(v.data[:, 0], v.data[:, 1]), rtol=1e-14)
                np.testing.assert_allclose(fast[1][row, : v.size], np.arctan2(v.data[:, 1], v.data[:, 0]), rtol=1e-13,
                                           atol=1e-15)
                np.testing.assert_array_equal(fast[2][row, : v.size], v.intensity)
        else:
            vecs = [v for pv in sim.coordinates for v in pv] if sim.has_multiple_phases else list(sim.coordinates)
            from diffsims_b200.simulations import get_closest
            for row, v in enumerate(vecs):
                rr_, tt_ = v.to_flat_polar()
                ri, ti = get_closest(kw["radial_axes"], rr_), get_closest(kw["azimuthal_axes"], tt_)
                m = (ri < len(kw["radial_axes"]) - 1) & (ti < len(kw["azimuthal_axes"]) - 1)
                np.testing.assert_array_equal(fast[0][row, : m.sum()], ri[m])
                np.testing.assert_array_equal(fast[1][row, : m.sum()], ti[m])
        for a, b in zip(fast, slow):
            assert a.shape == b.shape and a.dtype == b.dtype
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------- rotation-list producers on the device
def test_get_beam_directions_grid_matches_reference(golden_dir):
    from diffsims_b200.generators.rotation_list_generators import (beam_directions_device,
                                                                   get_beam_directions_grid,
                                                                   get_grid_around_beam_direction)
    gold = np.load(golden_dir / "beam_grid.npz")
    for system in cases.BEAM_GRID_SYSTEMS:   # test_rotation_list_generator.py:80-94
        grid = get_beam_directions_grid(system, 2)
        assert grid.shape == (cases.BEAM_GRID_SIZES_2DEG[system], 3)
        ref = K.beam_directions_grid(system, 2)
        np.testing.assert_allclose(grid, ref, rtol=0, atol=1e-10)   # same points, same order
    for key in gold.files:
        if key.endswith("_5deg"):
            mesh, system = key[:-5].rsplit("_", 1)
            np.testing.assert_allclose(get_beam_directions_grid(system, 5, mesh=mesh), gold[key], atol=1e-10)
    with pytest.raises(NotImplementedError):
        get_beam_directions_grid("cubic", 10, mesh="invalid")
    # the quaternions are the active form of Rotation.from_euler(grid)
    euler, quat = beam_directions_device("cubic", 1.0)
    ref_q = (~Rotation.from_euler(euler.cpu().numpy(), degrees=True)).data
    np.testing.assert_allclose(quat.cpu().numpy(), ref_q, atol=1e-12)
    # a >= 300k grid straight into HBM, usable by the simulate kernel
    euler, quat = beam_directions_device("cubic", 0.058, want_euler=False)
    assert euler is None and quat.shape[0] > 300_000
    np.testing.assert_allclose(quat.norm(dim=1).cpu().numpy(), 1.0, atol=1e-12)
    gen = ds.SimulationGenerator(200)
    gt = gen._g_table(cases.phase("si"), 1.0, True, {})
    spots = engine_simulate(gt, quat[:4096], gen)
    assert int(spots.count.min()) >= 2   # at least the (doubled) direct beam everywhere
    g = get_grid_around_beam_direction((0, 45, 30), 5)
    assert len(g) == 72 and all(len(t) == 3 for t in g)
    np.testing.assert_allclose([t[1] for t in g], 45.0, atol=1e-2)


@pytest.mark.parametrize("mesh", ["uv_sphere", "normalized_cube", "spherified_cube_edge", "spherified_cube_corner",
                                  "icosahedral", "random"])
def test_get_beam_directions_grid_all_meshes(mesh):
    """diffsims/tests/generators/test_rotation_list_generator.py:53-76 (every mesh x crystal system), checked
    against the oracle instead of only being executed."""
    from diffsims_b200.generators import sphere_mesh_generators as smg
    from diffsims_b200.generators.rotation_list_generators import _points_to_grid, get_beam_directions_grid
    for system in cases.BEAM_GRID_SYSTEMS:
        if mesh == "random":   # unseeded in the reference: same vertices into both implementations
            pts = smg.get_random_sphere_vertices(5, seed=17)
            got = _points_to_grid(pts, system, want_quaternions=False)[0].cpu().numpy()
            assert get_beam_directions_grid(system, 5, mesh=mesh).shape[1] == 3
            ref = K.beam_directions_grid(system, 5, mesh=mesh, points=pts)
        else:
            got = get_beam_directions_grid(system, 5, mesh=mesh)
            ref = K.beam_directions_grid(system, 5, mesh=mesh)
        assert got.shape == ref.shape and got.shape[0] > 0
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)


def test_beam_directions_grid_to_euler():
    """diffsims/tests/generators/test_sphere_mesh_generators.py:124-146."""
    from diffsims_b200.generators.sphere_mesh_generators import beam_directions_grid_to_euler
    grid = np.array([[1.0, 0, 0], [0, 1, 0], [0, 1, 1], [1, 0, 1]])
    grid = (grid.T / np.linalg.norm(grid, axis=1)).T
    np.testing.assert_allclose(beam_directions_grid_to_euler(grid), [[0, 90, 90], [0, 90, 0], [0, 45, 0], [0, 45, 90]],
                               atol=1e-12)
    pts = np.random.default_rng(4).normal(size=(1000, 3))
    pts[0] = (0, 0, 1)       # pole: arccos(0 / 0) -> nan_to_num
    pts[1] = (0, 0, -2)
    np.testing.assert_allclose(beam_directions_grid_to_euler(pts), K.beam_directions_grid_to_euler(pts), atol=1e-10)


def engine_simulate(gt, quat, gen):
    from diffsims_b200 import engine
    return engine.simulate(gt, quat, gen.wavelength, 0.01, 0.01, "lorentzian")


def test_polar_flatten_user_built_simulations():
    """diffsims/tests/simulations/test_simulations2d.py:101-127, :346-355, :472-480 (shapes, axis snapping)."""
    al = Phase(name="al", space_group=225,
               structure=Structure(atoms=[Atom("al", [0, 0, 0])], lattice=Lattice(0.405, 0.405, 0.405, 90, 90, 90)))
    gen = ds.SimulationGenerator(accelerating_voltage=200)
    coords = DiffractingVector(phase=al, xyz=[[1, 0, 0], [2, 0, 0], [3, 3, 0], [-4, 0, 0], [-5, 0, 0], [-6, 0, 0],
                                              [-7, 0, 0], [-8, 0, 0]], intensity=[1, 2, 3, 4, 5, 6, 7, 8])
    sim = Simulation2D(phases=al, simulation_generator=gen, coordinates=coords,
                       rotations=Rotation.from_euler([[0, 45, 0]], degrees=True))
    r, t, i = sim.polar_flatten_simulations()
    assert r.shape == t.shape == i.shape == (1, 8)
    np.testing.assert_allclose(r[0], np.hypot(coords.data[:, 0], coords.data[:, 1]))
    np.testing.assert_allclose(t[0], np.arctan2(coords.data[:, 1], coords.data[:, 0]))
    r, t, i = sim.polar_flatten_simulations(radial_axes=np.linspace(0, 7, 5), azimuthal_axes=np.linspace(0, 2 * np.pi, 10))
    assert r.shape == t.shape == i.shape == (1, 8) and r.dtype.kind == "i"
    np.testing.assert_array_equal(r[:, 6:], 0)    # the last two are beyond the radial axis
    np.testing.assert_array_equal(t[:, 6:], 0)
    np.testing.assert_array_equal(i[:, 6:], 0)
    rot4 = Rotation.from_euler([[0, a, 0] for a in (0, 15, 30, 45)], degrees=True)
    c4 = DiffractingVector(phase=al, xyz=[[1, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]], intensity=[1, 2, 3, 4])
    multi = Simulation2D(phases=al, simulation_generator=gen, coordinates=[c4] * 4, rotations=rot4)
    assert multi.polar_flatten_simulations()[0].shape == (4, 4)
    p2 = al.deepcopy()
    p2.name = "al2"
    mp = Simulation2D(phases=[al, p2], simulation_generator=gen, coordinates=[[c4] * 4, [c4] * 4], rotations=[rot4, rot4])
    assert mp.polar_flatten_simulations()[0].shape == (8, 4)


class TestGetPatternFromPixelCoordinatesAndIntensities:
    """diffsims/tests/patterns/test_detector_functions.py:160-300, plus the oracle on random spots."""

    @staticmethod
    def f(*a, **k):
        from diffsims_b200.pattern.detector_functions import get_pattern_from_pixel_coordinates_and_intensities
        return get_pattern_from_pixel_coordinates_and_intensities(*a, **k)

    def test_2d_vs_3d_coordinates(self):
        c2 = np.asarray([[10, 10], [20, 30], [15, 20]])
        c3 = np.asarray([[10, 10, 92], [20, 30, 0], [15, 20, -192]])
        assert np.array_equal(self.f(c2, np.ones(3) * 100, (50, 50), 1), self.f(c3, np.ones(3) * 100, (50, 50), 1))

    def test_integer_vs_float_coordinates(self):
        ci = np.asarray([[10, 10], [20, 30], [15, 20]]).astype(int)
        pi = self.f(ci, np.ones(3) * 100, (50, 50), 1).astype(int)
        pf = self.f(ci.astype(float), np.ones(3) * 100, (50, 50), 1).astype(int)
        assert np.allclose(pi, pf)

    def test_low_intensity(self):
        ci = np.asarray([[10, 10], [20, 30], [15, 20]]).astype(float)
        assert np.sum(self.f(ci, np.ones(3), (50, 50), 1)) == 0.0

    def test_total_intensity_preservation(self):
        p = self.f(np.asarray([[10, 10]]).astype(int), np.array([100]), (50, 50), 3)
        assert np.allclose(np.sum(p), 100)
        ci = np.asarray([[10, 10]]).astype(float)
        assert np.sum(self.f(ci, np.array([1000]), (50, 50), 1)) / 1000 > 0.999
        assert np.sum(self.f(ci, np.array([20]), (50, 50), 1, clip_threshold=0.01)) / 20 > 0.999

    def test_spot_in_corner(self):
        ci = np.asarray([[0.3, 0.1]])
        assert np.sum(self.f(ci, np.array([100]), (50, 50), 3)) < 100
        assert np.allclose(np.sum(self.f(ci.astype(int), np.array([100]), (50, 50), 3)), 100)

    @pytest.mark.parametrize("integer", [True, False])
    def test_matches_oracle(self, integer):
        rng = np.random.default_rng(5)
        shape = (70, 90)
        xy = rng.uniform(-6, 96, (40, 2))          # floats may lie outside the frame and still spread into it
        inten = rng.uniform(20, 900, 40)
        if integer:
            xy = np.stack([rng.integers(-90, 90, 40), rng.integers(-70, 70, 40)], axis=1)   # negative: numpy wrap
        got = self.f(xy, inten, shape, 2.5)
        ref = K.pattern_from_pixel_coordinates_and_intensities(xy, inten, shape, 2.5)
        assert got.dtype == np.float64 and got.shape == shape
        assert np.abs(got - ref).max() <= 1e-4 * ref.max()
        if integer:
            with pytest.raises(IndexError):
                self.f(np.array([[95, 3]]), np.array([1.0]), shape, 2.5)

    def test_spots_outside_the_frame_match_reference_golden(self, golden_dir):
        gold = np.load(golden_dir / "detector.npz")
        xy, xy_int, inten = cases.detector_spots_outside((70, 90), 40, 5)
        for coords, key in ((xy, "outside_float"), (xy_int, "outside_int")):
            got = self.f(coords, inten, (70, 90), 2.5)
            assert np.abs(got - gold[key]).max() <= 1e-4 * gold[key].max()


def test_library_builder_capacity_covers_the_pre_cut_count():
    """K2 writes every reflection that passes the excitation-error cut and compacts after the minimum-intensity cut,
    so the calibrated capacity must cover the count BEFORE that cut (with the extinction marking off, Si has five
    forbidden reflections per allowed one); unchecked passes expose the device-side maximum for a later check."""
    import torch
    from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
    from tests.helpers import random_quats
    gen = ds.SimulationGenerator(200, shape_factor_model="sinc")     # sinc: no extinction marking
    assert gen._extinct_rel_cut(True) == 0.0
    b = TemplateLibraryBuilder(gen, cases.phase("si"), reciprocal_radius=2.0, max_excitation_error=0.05,
                               shape=(64, 64), sigma=2, calibration=2 / 32)
    b.prepare()
    q = np.vstack([[1.0, 0, 0, 0], random_quats(63, 2)])
    qd = torch.as_tensor(active_quaternions(q), device=engine_device())
    cap = b.calibrate_cap(qd)
    sp = b.simulate(qd)
    assert int(sp.max_count.item()) <= cap and int(sp.count.max().item()) < int(sp.max_count.item())
    b.assert_no_overflow(sp)
    ref = gen.calculate_diffraction2d(cases.phase("si"), Rotation(q), reciprocal_radius=2.0, max_excitation_error=0.05)
    for r in (0, 1, 17):
        assert int(sp.count[r]) == ref.coordinates[r].size
    b.cap = 32                                   # too small on purpose: the check must notice
    sp = b.simulate(qd)
    with pytest.raises(RuntimeError):
        b.assert_no_overflow(sp)


def engine_device():
    from diffsims_b200 import engine
    return engine.device()


# --------------------------------------------------------------------------- round-2 closures
@pytest.mark.parametrize("tag,kw", [("aniso", {}), ("aniso_rot", dict(in_plane_angle=30.0, mirrored=True))])
def test_old_api_pattern_with_per_axis_calibration_matches_reference_golden(golden_dir, tag, kw):
    """DiffractionSimulation.get_diffraction_pattern with calibration = (cx, cy), diffsims/sims/diffraction_simulation.py
    :141-147, :296-354, against the reference executed by tests/golden/make_golden.py."""
    g = np.load(golden_dir / "closures.npz")
    sim = ds.DiffractionSimulation(g[f"{tag}_coords"], intensities=g[f"{tag}_intensities"],
                                   calibration=tuple(g[f"{tag}_calibration"]), with_direct_beam=True)
    got = sim.get_diffraction_pattern(shape=(256, 256), sigma=6, **kw)
    ref = g[f"{tag}_pattern"]
    assert got.shape == ref.shape and np.abs(got - ref).max() <= IMG_ATOL
    with pytest.raises(NotImplementedError):
        sim.get_diffraction_pattern(shape=(256, 200))


@pytest.mark.parametrize("angle", [0, 90, 45])
def test_knife_edge_pixels_are_decided_without_the_references_ulp_noise(golden_dir, angle):
    """Spots whose pixel coordinate is an exact integer (here: multiples of 25 px): the reference decides the truncation
    by the <= 1 ulp noise of r cos(atan2(y, x) + a) + cx (27.999999999999996 -> pixel 27), the kernel's algebraic form does
    not (28.0 -> pixel 28).  Characterised against the reference golden: every disagreement is a spot on such a knife edge,
    it moves by exactly one pixel, and the rendered template is the oracle's rasterisation of the kernel's pixel choice."""
    g = np.load(golden_dir / "closures.npz")
    coords, inten = g["knife_coords"], g["knife_intensities"]
    live = np.any(coords, axis=1)                       # with_direct_beam=False masks (0, 0, 0), :171-179
    ref_px = g[f"knife_pixels_{angle}"]
    a = np.deg2rad(angle)
    ca, sa = (1.0, 0.0) if angle == 0 else (np.cos(a), np.sin(a))
    xs, ys = coords[live, 0] / 0.01, coords[live, 1] / 0.01
    ours = np.stack([(xs * ca - ys * sa) + 128, (ys * ca + xs * sa) + 128], axis=1)   # the kernels' documented rule (no FMA)
    assert np.abs(ours - ref_px).max() < 1e-12           # the same numbers up to the reference's round-off ...
    differ = (ours.astype(int) != ref_px.astype(int)).any(axis=1)
    assert differ.sum() <= 0.25 * len(ours)              # ... (measured: 8 / 17 / 2 of 80 spots at 0 / 90 / 45 degrees)
    on_edge = np.abs(ref_px - np.rint(ref_px)).min(axis=1) < 1e-9
    assert np.all(on_edge[differ])                        # ... and only knife-edge spots can land one pixel apart
    assert np.abs(ours.astype(int) - ref_px.astype(int)).max() <= 1
    sim = ds.DiffractionSimulation(coords, intensities=1.0 + np.arange(len(coords)) % 7, calibration=0.01)
    got = sim.get_diffraction_pattern(shape=(256, 256), sigma=2, in_plane_angle=angle)
    inside = (ours[:, 0] >= 0) & (ours[:, 0] < 256) & (ours[:, 1] >= 0) & (ours[:, 1] < 256)
    expect = K.pattern_from_pixel_coordinates_and_intensities(ours[inside].astype(int), inten[inside], (256, 256), 2, 1)
    expect = expect / expect.max()
    assert np.abs(got - expect).max() <= IMG_ATOL
    if not differ.any():
        assert np.abs(got - g[f"knife_pattern_{angle}"]).max() <= IMG_ATOL


def test_objects_exposing_only_the_orix_and_diffpy_attribute_surface():
    """The drop-in takes orix.Phase / orix.Rotation / diffpy.structure objects by duck typing.  Neither library is in the
    image, so this feeds minimal objects that expose EXACTLY the attribute surface SURVEY.md section 8b lists (anything
    else raises AttributeError) through calculate_diffraction2d + get_diffraction_pattern: same result as the stand-ins."""
    import copy

    class OnlyThese:
        _allowed = ()

        def __getattr__(self, name):        # only called for names that are not real attributes
            raise AttributeError(f"{type(self).__name__} does not expose .{name} (not part of the orix / diffpy surface used)")

    class MinLattice(OnlyThese):            # diffpy.structure.Lattice: base, recbase, stdbase, baserot, rnorm, setLatPar, abcABG
        def __init__(self, lat):
            self.base, self.recbase = np.array(lat.base), np.array(lat.recbase)
            self.stdbase, self.baserot = np.array(lat.stdbase), np.array(lat.baserot)
            self._abc = lat.abcABG()

        def rnorm(self, hkl):
            return np.sqrt(((np.asarray(hkl, float) @ self.recbase.T) ** 2).sum(axis=-1))

        def abcABG(self):
            return self._abc

        def setLatPar(self, baserot=None, **kw):
            assert not kw
            old = self.baserot
            self.baserot = np.array(baserot)
            rot = np.linalg.solve(old, self.baserot)          # base = stdbase @ baserot
            self.base = self.base @ rot
            self.recbase = np.linalg.inv(self.base)

    class MinAtom(OnlyThese):               # diffpy.structure.Atom: element, xyz, occupancy
        def __init__(self, a):
            self.element, self.xyz, self.occupancy = a.element, np.array(a.xyz), a.occupancy

    class MinStructure(OnlyThese):          # diffpy.structure.Structure: iterable of atoms, .lattice
        def __init__(self, st):
            self._atoms = [MinAtom(a) for a in st]
            self.lattice = MinLattice(st.lattice)

        def __iter__(self):
            return iter(self._atoms)

        def __len__(self):
            return len(self._atoms)

    class MinPhase(OnlyThese):              # orix.crystal_map.Phase: name, structure, point_group, deepcopy()
        def __init__(self, ph):
            self.name, self.point_group, self.structure = ph.name, ph.point_group, MinStructure(ph.structure)

        def deepcopy(self):
            return copy.deepcopy(self)

    class MinRotation(OnlyThese):           # orix.quaternion.Rotation: data, size, to_matrix(), ~, iteration / indexing
        def __init__(self, data):
            self.data = np.atleast_2d(np.asarray(data, float))

        @property
        def size(self):
            return self.data.shape[0]

        @property
        def shape(self):
            return (self.data.shape[0],)

        def to_matrix(self):
            return Rotation(self.data).to_matrix()

        def __invert__(self):
            q = self.data.copy()
            q[:, 1:] *= -1
            return MinRotation(q)

        def __getitem__(self, key):
            return MinRotation(self.data[key])

        def __iter__(self):
            return (MinRotation(self.data[i]) for i in range(self.size))

        def __len__(self):
            return self.size

    phase = cases.phase("ti")               # hexagonal: exercises the a || x, c* || z realignment of the lattice
    rot = Rotation.random(5, rng=3)
    gen = ds.SimulationGenerator(300)
    ref = gen.calculate_diffraction2d(phase, rot, reciprocal_radius=1.5, max_excitation_error=0.02)
    got = gen.calculate_diffraction2d(MinPhase(phase), MinRotation(rot.data), reciprocal_radius=1.5, max_excitation_error=0.02)
    for i in range(5):
        a, b = ref.coordinates[i], got.coordinates[i]
        np.testing.assert_array_equal(a.hkl, b.hkl)
        np.testing.assert_array_equal(a.data, b.data)
        np.testing.assert_array_equal(a.intensity, b.intensity)
    kw = dict(shape=(128, 128), sigma=4, calibration=1.5 / 64)
    np.testing.assert_array_equal(ref.get_diffraction_pattern(**kw), got.get_diffraction_pattern(**kw))
    one = gen.calculate_diffraction2d(MinPhase(phase), MinRotation(rot.data[:1]), reciprocal_radius=1.5)
    assert one.coordinates.size > 0


def test_uint16_export_of_normalised_templates():
    """The optional 16-bit export (ds_quantize_u16 / run_host into uint16 buffers): rint(v * 65535) of the float32
    templates -- quantisation 0.5 / 65535 = 7.6e-6 of the peak, the maximum pixel stays exactly 65535."""
    import torch
    from diffsims_b200 import engine
    from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
    from tests.helpers import random_quats
    gen = ds.SimulationGenerator(200)
    b = TemplateLibraryBuilder(gen, make_phase(), reciprocal_radius=1.5, max_excitation_error=0.02, shape=(144, 200), sigma=4.0,
                               calibration=1.5 / 72)
    n = 37
    q = torch.as_tensor(active_quaternions(random_quats(n, 4))).pin_memory()
    f32 = torch.empty((n, 144, 200), dtype=torch.float32).pin_memory()
    u16 = torch.empty((n, 144, 200), dtype=torch.uint16).pin_memory()
    _, d2h_f = b.run_host(q, out_host=f32, chunk=16)
    _, d2h_u = b.run_host(q, ring=[torch.empty((16, 144, 200), dtype=torch.uint16).pin_memory() for _ in range(2)], chunk=16,
                          consumer=lambda lo, hi, view: u16[lo:hi].copy_(view))
    assert d2h_u * 2 == d2h_f
    ref = np.rint(np.clip(f32.numpy().astype(np.float64), 0, 1) * 65535)
    got = u16.numpy().astype(np.float64)
    assert np.abs(got - ref).max() <= 1          # (float32 product v * 65535 against the float64 one: ties only)
    assert np.abs(got / 65535 - f32.numpy()).max() <= 0.5 / 65535 + 1e-7
    assert all(got[i].max() == 65535 for i in range(n) if f32[i].max() > 0)
    # odd sizes go through the scalar tail
    x = torch.rand(1003, device=engine.device())
    assert torch.equal(engine.quantize_u16(x).cpu().to(torch.int32), torch.round(x * 65535).cpu().to(torch.int32))
