"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the host logic of the
API mirror (containers, crystal stand-ins, g-set enumeration) behaves like the reference's."""
import ctypes
import pickle
import re
from pathlib import Path

import numpy as np
import pytest

import diffsims_b200 as ds
from diffsims_b200 import _cabi, engine
from diffsims_b200.crystal import Atom, Lattice, Phase, Rotation, Structure
from diffsims_b200.crystallography import DiffractingVector, ReciprocalLatticeVector, g_set_from_min_dspacing
from diffsims_b200.simulations import Simulation2D
from diffsims_b200.utils import shape_factor_models as sfm
from oracle import kinematical as K
from tests.golden import cases

ROOT = Path(__file__).resolve().parents[1]


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "diffsims_b200.h").read_text()
    declared = set(re.findall(r"\b(ds_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS)
    lib = _cabi.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.ds_abi_version() == _cabi.ABI_VERSION
    assert isinstance(lib.ds_last_error(), bytes)


def test_abi_argument_validation_without_gpu():
    """Argument checks happen before any CUDA call, so they can be exercised on a CPU box."""
    lib = _cabi.lib()
    rc = lib.ds_simulate(None, 4, None, 10, None, None, None, 1.0, 40.0, 0.01, 0.01, 99, 5.0, 0.0, 1e-20, 32,
                         None, None, None, None, None, None, 0, None, None, None)
    assert rc != 0 and b"shape model" in lib.ds_last_error()
    rc = lib.ds_render(None, 1, 32, None, None, None, 64, 64, 0.0, 32.0, 32.0, 0.0, 0, 1, 2.0, 8, 1.0, 1, None, None, 0.0)
    assert rc != 0 and b"calibration" in lib.ds_last_error()
    rc = lib.ds_structure_factors(None, 4, None, None, 1, None, None, 1, None, None, None, 7, None, None, None, 0, None)
    assert rc != 0 and b"scattering" in lib.ds_last_error()
    assert lib.ds_set_option(b"no_such_option", 1) != 0 and b"unknown option" in lib.ds_last_error()
    # per template the larger of the two tcgen05 record layouts: 32 + 8 cap + 4 cap + 8 ceil(cap / 16) (per-reflection
    # kernel) and 32 + 8 cap + 528 (row-binned kernel: row offsets), 16-aligned
    assert lib.ds_render_scratch_bytes(1000, 32) == 16 + 1000 * 816
    assert lib.ds_render_scratch_bytes(1000, 992) == 16 + 1000 * 12432
    # phase tables | index box (int32, 16-aligned) | one result box (complex128) per atom unit (up to 16)
    assert lib.ds_structure_factors_scratch_bytes(500, 30) == 3 * 500 * 61 * 16 + (61 ** 3 * 4 + 15) // 16 * 16 + 16 * 61 ** 3 * 16
    assert lib.ds_structure_factors_scratch_bytes(40, 100) == 3 * 40 * 201 * 16      # H > 63: phase tables only
    assert lib.ds_render_launch_count(32, 256, 256, 40, 1, 3.8) == 1 and lib.ds_render_launch_count(288, 256, 256, 40, 1, 177.0) == 2
    assert lib.ds_render_launch_count(288, 512, 512, 40, 1, 177.0) == 1    # larger than a tensor-memory template: float32 kernels


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.device()
    gen = ds.SimulationGenerator(200)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gen.calculate_diffraction2d(cases.phase("si"))


def test_product_does_not_import_oracle():
    for f in (ROOT / "diffsims_b200").rglob("*.py"):
        assert "oracle" not in f.read_text(), f


# ------------------------------------------------------------------ crystal stand-ins
def test_lattice_conventions_match_diffpy():
    lat = Lattice(2.464, 2.464, 6.711, 90, 90, 120)
    # diffpy standard setting: a* || x, c || z
    np.testing.assert_allclose(lat.recbase[:, 0] / np.linalg.norm(lat.recbase[:, 0]), [1, 0, 0], atol=1e-15)
    np.testing.assert_allclose(lat.base[2], [0, 0, 6.711], atol=1e-15)
    np.testing.assert_allclose(lat.base @ lat.recbase, np.eye(3), atol=1e-15)
    np.testing.assert_allclose(lat.abcABG(), (2.464, 2.464, 6.711, 90, 90, 120))
    rec = lat.reciprocal()
    np.testing.assert_allclose(rec.base, lat.recbase.T)
    np.testing.assert_allclose(lat.rnorm([[1, 0, 0]]), rec.norm([[1, 0, 0]]))
    lat2 = Lattice(base=lat.base)
    np.testing.assert_allclose(lat2.abcABG(), lat.abcABG(), rtol=1e-12)
    np.testing.assert_allclose(lat2.stdbase @ lat2.baserot, lat.base, atol=1e-12)


def test_phase_realigns_like_orix():
    p = cases.phase("triclinic")
    lat = p.structure.lattice
    # a || x, c* || z
    np.testing.assert_allclose(lat.base[0] / np.linalg.norm(lat.base[0]), [1, 0, 0], atol=1e-12)
    cstar = lat.recbase[:, 2]
    np.testing.assert_allclose(cstar / np.linalg.norm(cstar), [0, 0, 1], atol=1e-12)
    # the reference's frame change (sim_utils.py:290-291) recovers the user's fractional coordinates
    mat = np.linalg.inv(lat.stdbase @ lat.recbase)
    np.testing.assert_allclose(p.structure.xyz @ mat, cases.structure("triclinic").xyz, atol=1e-12)
    assert p.deepcopy() is not p


def test_rotation_conventions_match_orix():
    e = np.deg2rad([[10, 20, 30], [200, 95, 300]])
    r = Rotation.from_euler(e)
    for i in range(2):
        np.testing.assert_allclose(r.to_matrix()[i], K.bunge_matrix(*e[i]), atol=1e-15)
        np.testing.assert_allclose(K.euler2mat_rzxz(*e[i]), K.bunge_matrix(*e[i]).T, atol=1e-15)
    np.testing.assert_allclose(r.to_euler(), e, atol=1e-12)
    np.testing.assert_allclose((~r).to_matrix(), np.transpose(r.to_matrix(), (0, 2, 1)), atol=1e-15)
    np.testing.assert_allclose(Rotation.from_matrix(r.to_matrix()).data, r.data, atol=1e-12)
    np.testing.assert_allclose((r * ~r).data, [[1, 0, 0, 0]] * 2, atol=1e-15)
    assert r.size == 2 and r[0].size == 1 and len(list(r)) == 2
    assert Rotation.random(5).size == 5


def test_g_set_enumeration():
    al = Phase("al", space_group=225, structure=Structure([Atom("Al", [0, 0, 1])],
                                                          Lattice(4.04, 4.04, 4.04, 90, 90, 90)))
    hkl = g_set_from_min_dspacing(al.structure.lattice, 0.7)
    assert len(hkl) == 798 and tuple(hkl[0]) == (5, 2, 2) and tuple(hkl[-1]) == (-5, -2, -2)
    rlv = ReciprocalLatticeVector.from_min_dspacing(al, 1.0, include_zero_vector=True)
    assert rlv.size == 257 and tuple(rlv.hkl[-1]) == (0, 0, 0)
    for name in cases.STRUCTURES:
        lat = cases.phase(name).structure.lattice
        np.testing.assert_array_equal(g_set_from_min_dspacing(lat, 1 / 1.3, True),
                                      K.from_min_dspacing(lat, 1 / 1.3, True)[0])


def test_reciprocal_lattice_vector_basics():
    al = Phase("al", space_group=225, structure=Structure([Atom("Al", [0, 0, 1])],
                                                          Lattice(4.04, 4.04, 4.04, 90, 90, 90)))
    rlv = DiffractingVector(al, hkl=[[1, 1, 1], [2, 0, 0]])
    np.testing.assert_allclose(rlv.gspacing, [0.42872545, 0.4950495], rtol=1e-7)
    np.testing.assert_allclose(rlv.hkl, [[1, 1, 1], [2, 0, 0]], atol=1e-12)
    assert np.isnan(rlv.intensity).all()
    rlv.intensity = 3
    np.testing.assert_array_equal(rlv.intensity, [3, 3])
    with pytest.raises(ValueError):
        rlv.intensity = [1, 2, 3]
    with pytest.raises(ValueError):
        DiffractingVector(al, hkl=[[1, 1, 1]], intensity=[1, 2])
    with pytest.raises(ValueError):
        ReciprocalLatticeVector(al)
    with pytest.raises(NotImplementedError):
        rlv.calculate_structure_factor()
    assert rlv[0].size == 1 and rlv[0].intensity[0] == 3
    rot = Rotation.from_euler([[0, 90, 90]], degrees=True)
    rr = rlv.rotate_with_basis(rot)
    np.testing.assert_allclose(rr.data, rlv.data @ rot.to_matrix()[0], atol=1e-15)
    np.testing.assert_allclose(rr.hkl, rlv.hkl, atol=1e-12)   # indices survive the basis rotation
    with pytest.raises(ValueError):
        rlv.rotate_with_basis(Rotation.random(2))
    r, t = rlv.to_flat_polar()
    np.testing.assert_allclose(r, np.hypot(rlv.data[:, 0], rlv.data[:, 1]))
    # diffsims/tests/crystallography/test_diffracting_vector.py:10-18, :24-40, :47-75
    fe = Phase("ferrite", space_group=229, structure=Structure([Atom("Fe", [0, 0, 0]), Atom("Fe", [.5, .5, .5])],
                                                               Lattice(2.8665, 2.8665, 2.8665, 90, 90, 90)))
    dv = DiffractingVector(fe, hkl=[[1, 1, 1], [2, 0, 0]], intensity=[1, 2])
    assert dv.phase == fe and dv.shape == (2,) and dv.hkl.shape == (2, 3)
    np.testing.assert_allclose(dv.intensity, [1, 2])
    np.testing.assert_allclose(dv.basis_rotation.to_matrix(), np.eye(3)[None], atol=1e-15)
    full = DiffractingVector.from_min_dspacing(fe, 1.5)
    full.intensity = 1
    assert isinstance(full.intensity, np.ndarray) and full[0:3].size == 3
    np.testing.assert_allclose(full[0:3].intensity, np.ones(3))
    with pytest.raises(ValueError):
        full.intensity = [0, 1]
    rot = Rotation.from_euler([[90, 90, 0]], degrees=True)
    fe2 = fe.deepcopy()
    fe2.structure.lattice.setLatPar(baserot=rot.to_matrix()[0])
    plain = ReciprocalLatticeVector(fe, hkl=[[1, 1, 1], [2, 0, 0]])
    np.testing.assert_allclose(DiffractingVector(fe2, xyz=plain.data @ rot.to_matrix()[0]).hkl, plain.hkl, atol=1e-12)
    r, t = DiffractingVector(fe, xyz=[[1, 1, 1], [0.5, -0.5, 0]]).to_flat_polar()
    np.testing.assert_allclose(r, [np.sqrt(2), 0.70710678])
    np.testing.assert_allclose(t, [np.pi / 4, -np.pi / 4])
    r, t = DiffractingVector(fe, xyz=[[[1, 1, 1], [0.5, -0.5, 0]], [[1, 1, 1], [0.5, -0.5, 0]]]).to_flat_polar()
    np.testing.assert_allclose(r, [np.sqrt(2), np.sqrt(2), 0.70710678, 0.70710678])
    np.testing.assert_allclose(t, [np.pi / 4, np.pi / 4, -np.pi / 4, -np.pi / 4])


# ------------------------------------------------------------------ Simulation2D container
@pytest.fixture
def al_phase():
    return Phase(name="al", space_group=225,
                 structure=Structure(atoms=[Atom("al", [0, 0, 0])], lattice=Lattice(0.405, 0.405, 0.405, 90, 90, 90)))


def _coords(al_phase, n=4):
    return DiffractingVector(phase=al_phase, xyz=[[1, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]][:n],
                             intensity=[1, 2, 3, 4][:n])


def test_simulation2d_single(al_phase):
    gen = ds.SimulationGenerator(accelerating_voltage=200)
    rot = Rotation.from_euler([[0, 45, 0]], degrees=True)
    sim = Simulation2D(phases=al_phase, simulation_generator=gen, coordinates=_coords(al_phase), rotations=rot)
    rotation, phase, coords = sim.get_simulation(0)
    assert phase == 0 and rotation.size == 1
    with pytest.raises(ValueError):
        sim.iphase[0]
    with pytest.raises(ValueError):
        sim.irot[0]
    assert sum(1 for _ in sim) == 1
    assert sim._num_rotations() == 1
    assert sim.deepcopy() is not sim
    with pytest.raises(ValueError):
        Simulation2D(phases=al_phase, simulation_generator=gen, coordinates=[_coords(al_phase)] * 2, rotations=rot)


def test_simulation2d_multi_rotation(al_phase):
    gen = ds.SimulationGenerator(accelerating_voltage=200)
    rot = Rotation.from_euler([[0, a, 0] for a in (0, 15, 30, 45)], degrees=True)
    c = _coords(al_phase)
    sim = Simulation2D(phases=al_phase, simulation_generator=gen, coordinates=[c, c, c, c], rotations=rot)
    assert isinstance(sim.coordinates, np.ndarray) and sim.current_size == 4
    for i in range(4):
        rotation, phase, coords = sim.get_simulation(i)
        assert phase == 0 and isinstance(coords, DiffractingVector)
    np.testing.assert_array_equal(sim.get_current_rotation_matrix(), rot[0].to_matrix()[0])
    assert sim.irot[0].rotations.size == 1 and sim.irot[0].coordinates.size == 4
    assert sim.irot[0:2].rotations.size == 2 and sim.irot[0:2].coordinates.size == 2
    assert sum(1 for _ in sim) == 4
    with pytest.raises(ValueError):
        Simulation2D(phases=al_phase, simulation_generator=gen, coordinates=[c, c, c], rotations=rot)


def test_simulation2d_multi_phase(al_phase):
    gen = ds.SimulationGenerator(accelerating_voltage=200)
    rot = Rotation.from_euler([[0, a, 0] for a in (0, 15, 30, 45)], degrees=True)
    c = _coords(al_phase)
    p2 = al_phase.deepcopy()
    p2.name = "al2"
    sim = Simulation2D(phases=[al_phase, p2], simulation_generator=gen, coordinates=[[c] * 4, [c] * 4],
                       rotations=[rot, rot])
    assert isinstance(sim.phases, np.ndarray) and isinstance(sim.rotations, np.ndarray)
    assert [sim.get_simulation(i)[1] for i in range(8)] == [0] * 4 + [1] * 4
    assert sim.iphase[0].rotations.size == 4 and sim.iphase["al2"].phases.name == "al2"
    with pytest.raises(ValueError):
        sim.iphase[3.1]
    assert sim.irot[0].rotations.size == 2 and sim.irot[0:2].rotations.size == 2
    assert sum(1 for _ in sim) == 8
    with pytest.raises(ValueError):
        Simulation2D(phases=[al_phase, p2], simulation_generator=gen, coordinates=[[c] * 2, [c] * 2],
                     rotations=[rot, rot])
    with pytest.raises(ValueError):
        Simulation2D(phases=[al_phase, p2], simulation_generator=gen, coordinates=[[c] * 4] * 3,
                     rotations=[rot, rot])
    with pytest.raises(ValueError):
        Simulation2D(phases=[al_phase, p2], simulation_generator=gen, coordinates=[[c] * 4] * 3,
                     rotations=[rot, rot, rot])


# ------------------------------------------------------------------ generator front end / misc host logic
def test_generator_init_and_errors():
    g = ds.SimulationGenerator(300)
    assert g.scattering_params == "lobato" and g.precession_angle == 0 and g.minimum_intensity == 1e-20
    assert g.shape_factor_model == sfm.lorentzian and g.approximate_precession is True
    assert repr(g) == ("SimulationGenerator(accelerating_voltage=300, scattering_params=lobato, "
                       "approximate_precession=True)")
    assert ds.SimulationGenerator(300, shape_factor_model="linear").shape_factor_model == sfm.linear
    assert ds.SimulationGenerator(300, shape_factor_model=sfm.binary).shape_factor_model == sfm.binary
    with pytest.raises(NotImplementedError):
        ds.SimulationGenerator(300, scattering_params="_empty")
    with pytest.raises(NotImplementedError):
        ds.SimulationGenerator(300, shape_factor_model="dracula")
    with pytest.raises(NotImplementedError):
        ds.DiffractionGenerator(300, scattering_params="_empty")
    np.testing.assert_almost_equal(g.wavelength, 0.0196874888)
    si = cases.phase("si")
    with pytest.raises(ValueError):   # phase / rotation count mismatch is checked before any device work
        g.calculate_diffraction2d([si, si], rotation=[Rotation.random(2)])


def test_shape_factor_callables_match_oracle():
    s = np.concatenate([np.linspace(-0.05, 0.05, 21), [0.0]])
    for name in ("linear", "sinc", "sin2c", "atanc", "lorentzian"):
        np.testing.assert_allclose(getattr(sfm, name)(s.copy(), 0.02), K.SHAPE_FACTOR_MODELS[name](s.copy(), 0.02),
                                   rtol=1e-13, atol=1e-15)
    assert sfm.binary(s, 0.02) == 1
    np.testing.assert_allclose(sfm.linear(0.5, 1), 0.5)


def test_atom_arrays_grouping_and_unknown_element():
    st = cases.structure("triclinic")
    frac, occ, start, coeffs, dw = engine.atom_arrays(st, cases.DW, "lobato")
    assert list(start) == [0, 1, 3, 4]            # Si | O, O | Fe  ("Si4+" and "O2-" lose their charge)
    np.testing.assert_allclose(occ, [1.0, 0.5, 1.0, 0.75])
    np.testing.assert_allclose(dw, [0.5, 0.7, 0.35])
    np.testing.assert_allclose(frac, st.xyz[[0, 1, 3, 2]], atol=1e-12)
    bad = Structure([Atom("Zz", [0, 0, 0])], Lattice(3, 3, 3, 90, 90, 90))
    with pytest.warns(UserWarning, match="not found in scattering parameter library"):
        _, _, _, c, _ = engine.atom_arrays(bad, {}, "xtables")
    assert not c.any()
    with pytest.raises(NotImplementedError):
        engine.get_scattering_params_dict("nope")


def test_estimate_cap_is_padded_and_bounded():
    assert engine.estimate_cap(690, 1.0, 0.01) % 32 == 0
    assert engine.estimate_cap(10, 1.0, 0.01) == 32
    assert engine.gaussian_radius(10) == 40 and engine.gaussian_radius(1.4) == 6


def test_old_api_containers(tmp_path):
    sim = ds.DiffractionSimulation(np.array([[0.0, 0, 0], [1, 2, 0], [3, 4, 0]]),
                                   indices=np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]]),
                                   intensities=np.array([9.0, 1, 2]), with_direct_beam=False)
    assert len(sim) == 2 and sim.size == 2                    # the (000) row is masked
    np.testing.assert_array_equal(sim.intensities, [1, 2])
    sim.calibration = 0.5
    np.testing.assert_allclose(sim.calibrated_coordinates, [[2, 4], [6, 8]])
    with pytest.raises(ValueError):
        sim.calibration = 0
    with pytest.raises(ValueError):
        ds.DiffractionSimulation(np.zeros((2, 3)), indices=np.zeros((3, 3)), intensities=np.zeros(2))
    assert sim[0].size == 1
    lib = ds.DiffractionLibrary()
    sims, px, ints = (np.empty(1, dtype=object) for _ in range(3))
    sims[0], px[0], ints[0] = sim, np.zeros((2, 2), int), sim.intensities
    lib["A"] = {"simulations": sims, "orientations": [(0, 0, 0)], "pixel_coords": px, "intensities": ints}
    e = lib.get_library_entry(phase="A", angle=(0, 0, 0))
    np.testing.assert_allclose(e["pattern_norm"], np.sqrt(5))
    with pytest.raises(ValueError):
        lib.get_library_entry(angle=(0, 0, 0))
    with pytest.raises(ValueError):
        lib.get_library_entry(phase="A", angle=(1, 1, 1))
    f = tmp_path / "lib.pkl"
    lib.pickle_library(f)
    with pytest.raises(RuntimeError):
        ds.libraries.load_DiffractionLibrary(f)
    back = ds.libraries.load_DiffractionLibrary(f, safety=True)
    np.testing.assert_array_equal(back["A"]["intensities"][0], [1, 2])
    with pytest.raises(ValueError):
        ds.StructureLibrary(["a"], [1, 2], [[(0, 0, 0)]])
    sl = ds.StructureLibrary(["a", "b"], [1, 2], [[(0, 0, 0)], [(0, 0, 0), (1, 1, 1)]])
    assert sl.get_library_size() == 3


# ---------------------------------------------------------------- sphere meshes (host vertex lists)
def test_sphere_mesh_vertices_match_reference(golden_dir):
    """Vertex lists of the uv-sphere / icosahedral / random / cube meshes: bit-identical to the reference's
    (golden arrays from sphere_mesh_generators.py) and the sizes its own tests pin
    (diffsims/tests/generators/test_sphere_mesh_generators.py:34-75, :115-121)."""
    from diffsims_b200.generators import sphere_mesh_generators as smg
    from oracle import kinematical as K
    gold = np.load(golden_dir / "beam_grid.npz")
    np.testing.assert_array_equal(smg.get_uv_sphere_mesh_vertices(7), gold["vertices_uv_sphere_7deg"])
    np.testing.assert_array_equal(smg.get_icosahedral_mesh_vertices(9), gold["vertices_icosahedral_9deg"])
    np.testing.assert_array_equal(smg.get_icosahedral_mesh_vertices(3), gold["vertices_icosahedral_3deg"])
    np.testing.assert_array_equal(smg.get_random_sphere_vertices(4, seed=3), gold["vertices_random_4deg_seed3"])
    assert smg.get_random_sphere_vertices(1).shape == (10313, 3)
    assert not np.allclose(smg.get_random_sphere_vertices(3, seed=7), smg.get_random_sphere_vertices(3, seed=8))
    for make, n in ((lambda: smg.get_uv_sphere_mesh_vertices(10), 614),
                    (lambda: smg.get_icosahedral_mesh_vertices(10), 642),
                    (lambda: smg.get_cube_mesh_vertices(10, "normalized"), 866),
                    (lambda: smg.get_cube_mesh_vertices(10, "spherified_edge"), 602),
                    (lambda: smg.get_cube_mesh_vertices(10, "spherified_corner"), 866)):
        grid = make()
        assert grid.shape == (n, 3)
        np.testing.assert_almost_equal(np.sum(grid), 0)
        assert np.unique(grid, axis=0).shape[0] == n
    for grid_type in ("normalized", "spherified_edge", "spherified_corner"):
        np.testing.assert_array_equal(smg.get_cube_mesh_vertices(6, grid_type), K.cube_mesh_vertices(6, grid_type))
    with pytest.raises(Exception):
        smg.get_cube_mesh_vertices(10, "non_existant")
    with pytest.raises(ValueError):
        smg.get_icosahedral_mesh_vertices(90)
    assert smg._max_neighbour_angle(np.array([[1.0, 0, 0], [0, 1, 0], [0, 1, 1], [1, 0, 1]])) == pytest.approx(45.0)  # :84-112


def test_header_and_ctypes_signatures_agree():
    """Every entry point declared in include/diffsims_b200.h is bound with the same number of arguments."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    header = re.sub(r"/\*.*?\*/", "", (root / "include" / "diffsims_b200.h").read_text(), flags=re.S)
    binding = (root / "diffsims_b200" / "_cabi.py").read_text()
    decls = re.findall(r"\b(?:int|int64_t|int32_t|const char \*)\s*\*?\s*(ds_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    assert {n for n, _ in decls} == set(_cabi.SYMBOLS)
    for name, args in decls:
        n = 0 if args.strip() in ("", "void") else len([a for a in args.split(",") if a.strip()])
        m = re.search(rf"L\.{name}\.argtypes = \[(.*?)\]", binding)
        if n == 0:
            assert m is None
        else:
            assert m is not None and len(m.group(1).split(",")) == n, name
