"""Parity of the three CUDA kernels (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from oracle import kinematical as K
from tests.golden import cases
from tests.helpers import (IMG_ATOL, RTOL, compare_spots, oracle_G_from_active_quat, random_quats, row)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffsims_b200 import engine
    return engine


# --------------------------------------------------------------------------- K1
@pytest.mark.parametrize("name", ["si", "graphite", "fe3c", "triclinic", "large"])
@pytest.mark.parametrize("sp", ["lobato", "xtables", None])
@pytest.mark.parametrize("aligned", [False, True])
def test_structure_factors(eng, name, sp, aligned):
    st = cases.phase(name).structure if aligned else cases.structure(name)
    hkl = cases.hkl_box(4 if name != "large" else 3)
    g = st.lattice.rnorm(hkl)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = K.kinematical_structure_factor(st, hkl, g, cases.DW, sp)
        F, I = eng.structure_factors(st, hkl, g, cases.DW, sp)
    F = F.cpu().numpy()
    got = F[:, 0] + 1j * F[:, 1]
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-11 * scale)
    np.testing.assert_allclose(I.cpu().numpy(), np.abs(ref) ** 2, rtol=1e-9, atol=1e-20 * scale ** 2)


@pytest.mark.parametrize("sp", ["lobato", "xtables", None])
@pytest.mark.parametrize("sparse", [False, True])
def test_structure_factors_factorised_kernels_match_the_direct_kernel(eng, sp, sparse):
    """Large cells with integer indices take the factorised kernels (phase tables per atom and axis; the box kernel
    register-tiles the index box, the row kernel handles sparse tables): both agree with the direct sincospi kernel to
    float64 round-off, rows that share an index triple (the second (000) of the new API) included."""
    import torch
    from diffsims_b200 import _cabi
    st = cases.phase("large").structure
    gs = K.GSet(st, 1.2, True)
    hkl = np.vstack([gs.hkl_int, [0, 0, 0], [1, -2, 3]])
    if sparse:  # a sparse table in a big box: every other row, and one far index -> the row kernel
        hkl = np.vstack([hkl[::2], [40, 0, 0]])
    assert len(hkl) >= 4096
    g = st.lattice.rnorm(hkl)
    dev = eng.device()
    atoms = eng.AtomTable(st, cases.DW, sp, dev)
    hkl_d, g_d = torch.as_tensor(hkl.astype(float), device=dev), torch.as_tensor(g, device=dev)
    pre = torch.as_tensor(np.linspace(0.5, 2.0, len(hkl)), device=dev)
    out = {}
    for name, H in (("direct", 0), ("factorised", int(np.abs(hkl).max()))):
        F = torch.zeros((len(hkl), 2), dtype=torch.float64, device=dev)
        I = torch.zeros((len(hkl),), dtype=torch.float64, device=dev)
        scratch = None
        if H:
            scratch = torch.empty(int(_cabi.lib().ds_structure_factors_scratch_bytes(atoms.n_atoms, H)), dtype=torch.uint8, device=dev)
        eng.launch_structure_factors(atoms, hkl_d, g_d, pre, F, I, hkl_int_max=H, scratch=scratch)
        out[name] = (F.cpu().numpy(), I.cpu().numpy())
    scale = np.abs(out["direct"][0]).max()
    np.testing.assert_allclose(out["factorised"][0], out["direct"][0], rtol=1e-9, atol=1e-11 * scale)
    np.testing.assert_allclose(out["factorised"][1], out["direct"][1], rtol=1e-9, atol=1e-20 * scale ** 2)
    assert not np.array_equal(out["factorised"][0], out["direct"][0])   # (a different kernel did run)


# --------------------------------------------------------------------------- K2
def _gtable_new_api(eng, phase, rr, with_direct_beam, sp="lobato", dw=None):
    st = phase.structure
    gs = K.GSet(st, rr, with_direct_beam)
    hkl, xyz = gs.hkl_int, gs.hkl_int @ st.lattice.recbase.T
    if with_direct_beam:  # the reference appends a second (000), simulation_generator.py:351-353
        hkl = np.vstack([hkl, [0, 0, 0]])
        xyz = np.vstack([xyz, [0, 0, 0]])
    return gs, eng.make_gtable(st, hkl, xyz, dw, sp)


@pytest.mark.parametrize("name,kv,rr,s_max,model,db", [
    ("si", 200, 1.0, 0.01, "lorentzian", True),
    ("si", 200, 2.0, 0.05, "lorentzian", True),
    ("si", 300, 5.0, 0.01, "lorentzian", True),
    ("al", 200, 1.0, 0.01, "lorentzian", True),
    ("ti", 300, 1.5, 0.02, "lorentzian", False),
    ("graphite", 200, 1.6768, 0.1, "lorentzian", False),
    ("fe3c", 200, 1.2, 0.03, "linear", True),
    ("fe3c", 200, 1.2, 0.03, "sinc", True),
    ("triclinic", 120, 1.0, 0.02, "sin2c", True),
    ("triclinic", 120, 1.0, 0.02, "atanc", False),
    ("fe_bcc", 200, 2.0, 0.01, "binary", True),
])
def test_simulate_matches_oracle(eng, name, kv, rr, s_max, model, db):
    phase = cases.phase(name)
    gs, gt = _gtable_new_api(eng, phase, rr, db)
    wl = K.get_electron_wavelength(kv)
    q = random_quats(24, 5)
    q[0] = (1, 0, 0, 0)                                     # zone axis
    q[1] = (np.cos(np.pi / 8), np.sin(np.pi / 8), 0, 0)     # 45 deg about x
    spots = eng.simulate(gt, q, wl, s_max, s_max, model, want_exc=True)
    total = 0
    for r in range(len(q)):
        ref = K.simulate_rotation(phase.structure, gs, oracle_G_from_active_quat(q[r]), wl, s_max,
                                  K.SHAPE_FACTOR_MODELS[model])
        total += compare_spots(ref, row(spots, r), s_max, rr)
    assert total > 0


def test_simulate_precession_approx(eng):
    phase = cases.phase("si")
    gs, gt = _gtable_new_api(eng, phase, 2.0, True)
    wl = K.get_electron_wavelength(300)
    q = random_quats(8, 9)
    q[0] = (1, 0, 0, 0)
    prec = np.deg2rad(0.5)
    spots = eng.simulate(gt, q, wl, 0.01, 0.01, "lorentzian_precession", precession_rad=prec, want_exc=True)
    for r in range(len(q)):
        ref = K.simulate_rotation(phase.structure, gs, oracle_G_from_active_quat(q[r]), wl, 0.01,
                                  precession_angle=0.5, approximate_precession=True)
        compare_spots(ref, row(spots, r), 0.01, 2.0, prec=True)


def test_simulate_reference_counts(eng):
    """Si, 300 kV, rr = 5, [001]: 70 reflections incl. the duplicated (000); 250 with precession
    (diffsims/tests/generators/test_simulation_generator.py:137-161)."""
    phase = cases.phase("si")
    gs, gt = _gtable_new_api(eng, phase, 5.0, True)
    wl = K.get_electron_wavelength(300)
    q = np.array([[1.0, 0, 0, 0]])
    spots = eng.simulate(gt, q, wl, 0.01, 0.01, "lorentzian")
    assert int(spots.count[0]) == 70
    spots = eng.simulate(gt, q, wl, 0.01, 0.01, "lorentzian_precession", precession_rad=np.deg2rad(0.5))
    assert int(spots.count[0]) == 250


@pytest.mark.parametrize("cta", [0, 1])
def test_simulate_streaming_tiles_large_cell(eng, opts, cta):
    """N_g > 6144: the double-buffered cp.async.bulk tile path of the warp-per-rotation kernel (sim_cta = 0) and the
    CTA-per-rotation kernel that large tables take by default, both against the oracle."""
    opts(sim_cta=cta)
    phase = cases.phase("large")
    gs, gt = _gtable_new_api(eng, phase, 1.2, True, dw=cases.DW)
    assert gt.n > 6144
    wl = K.get_electron_wavelength(200)
    q = random_quats(3, 11)
    spots = eng.simulate(gt, q, wl, 0.01, 0.01, "lorentzian", want_exc=True)
    for r in range(len(q)):
        ref = K.simulate_rotation(phase.structure, gs, oracle_G_from_active_quat(q[r]), wl, 0.01,
                                  debye_waller_factors=cases.DW)
        compare_spots(ref, row(spots, r), 0.01, 1.2)


@pytest.mark.parametrize("name,rr,s_max,model,prec_deg", [
    ("si", 2.0, 0.01, "lorentzian", 0.0),
    ("si", 1.0, 0.05, "linear", 0.0),
    ("triclinic", 1.5, 0.02, "sinc", 0.0),
    ("graphite", 2.0, 0.1, "lorentzian", 0.0),
    ("si", 2.0, 0.01, "lorentzian_precession", 0.5),
    ("large", 1.2, 0.01, "lorentzian", 0.0),
])
def test_simulate_scan_line_cull_is_identical(eng, opts, name, rr, s_max, model, prec_deg):
    """The scan-line cull (lattice lines solved for their slab crossing) and the brute-force cull feed the same
    candidates in the same order to the same float64 refine: identical reflection lists, including for lines
    parallel to the slab (c* in the detector plane) and zone axes."""
    phase = cases.phase(name)
    gs, gt = _gtable_new_api(eng, phase, rr, True)
    assert gt.n_lines > 0
    wl = K.get_electron_wavelength(200)
    q = random_quats(64, 21)
    q[0] = (1, 0, 0, 0)
    q[1] = (np.cos(np.pi / 4), np.sin(np.pi / 4), 0, 0)     # c* perpendicular to the beam: parallel lines
    q[2] = (np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0)
    q[3] = (np.cos(np.pi / 4 + 2e-4), np.sin(np.pi / 4 + 2e-4), 0, 0)   # almost parallel
    q[4] = (np.cos(np.pi / 4 + 2e-3), np.sin(np.pi / 4 + 2e-3), 0, 0)
    out = []
    for mode in (0, 1):
        opts(sim_lines=mode)
        sp = eng.simulate(gt, q, wl, s_max, s_max, model, precession_rad=np.deg2rad(prec_deg), want_exc=True)
        out.append(sp)
    a, b = out
    assert np.array_equal(a.count.cpu().numpy(), b.count.cpu().numpy())
    assert int(a.count.sum()) > 0
    for r in range(len(q)):
        n = int(a.count[r])
        assert np.array_equal(a.g_index[r, :n].cpu().numpy(), b.g_index[r, :n].cpu().numpy()), r
        for field in ("xyz", "intensity", "exc"):  # two template instantiations: fma contraction may differ
            x, y = getattr(a, field)[r, :n].cpu().numpy(), getattr(b, field)[r, :n].cpu().numpy()
            assert np.allclose(x, y, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(x).max())), (r, field)


@pytest.mark.parametrize("name,rr,s_max,model,prec_deg,min_int", [
    ("large", 1.2, 0.01, "lorentzian", 0.0, 1e-20),
    ("large", 1.2, 0.02, "linear", 0.0, 1e-3),
    ("si", 2.0, 0.05, "sinc", 0.0, 1e-20),
    ("graphite", 2.0, 0.1, "binary", 0.0, -1.0),
    ("si", 2.0, 0.01, "lorentzian_precession", 0.5, 1e-20),
    ("si", 1.5, 0.02, "sin2c", 0.3, 1e-20),      # numerically averaged precession: warp-cooperative evaluation
    ("triclinic", 1.5, 0.02, "return_s", 0.0, 1e-20),  # excitation errors returned, no intensity cut
])
def test_simulate_cta_per_rotation_is_identical(eng, opts, name, rr, s_max, model, prec_deg, min_int):
    """simulate_cta_kernel (few rotations over a large table: one CTA per rotation, the table split across its warps)
    returns exactly the reflection lists of the one-warp-per-rotation kernel -- with its default candidate stash, with a
    pool of a few chunks (some rotations fit, some do not) and through its fall-back alone (a single 64-entry chunk)."""
    phase = cases.phase(name)
    gs, gt = _gtable_new_api(eng, phase, rr, True)
    wl = K.get_electron_wavelength(200)
    q = random_quats(37, 5)
    q[0] = (1, 0, 0, 0)
    q[1] = (np.cos(np.pi / 4), np.sin(np.pi / 4), 0, 0)      # c* perpendicular to the beam: lines parallel to the slab
    q[2] = (np.cos(np.pi / 4 + 2e-4), np.sin(np.pi / 4 + 2e-4), 0, 0)   # almost parallel
    kw = dict(precession_rad=np.deg2rad(prec_deg), want_exc=True, min_intensity=min_int)
    opts(sim_lines=0, sim_cta=0)
    ref = eng.simulate(gt, q, wl, s_max, s_max, model, **kw)
    assert int(ref.count.sum()) > 0
    # (brute-force scan of the table slices; the interval expansion over lattice lines; small pools; the fall-back)
    for cta, stash, lines in ((1, -1, 0), (1, -1, 1), (1, 512, 0), (1, 512, 1), (1, 64, 0)):
        if lines and not gt.n_lines:
            continue
        opts(sim_cta=cta, sim_stash=stash, sim_lines=lines)
        got = eng.simulate(gt, q, wl, s_max, s_max, model, cap=ref.cap, **kw)
        assert int(got.max_count) == int(ref.max_count), (cta, stash)
        assert np.array_equal(got.count.cpu().numpy(), ref.count.cpu().numpy()), (cta, stash)
        for r in range(len(q)):
            n = int(ref.count[r])
            assert np.array_equal(got.g_index[r, :n].cpu().numpy(), ref.g_index[r, :n].cpu().numpy()), (cta, stash, r)
            for field in ("xyz", "intensity", "exc"):  # separate instantiations: fma contraction may differ
                x, y = getattr(ref, field)[r, :n].cpu().numpy(), getattr(got, field)[r, :n].cpu().numpy()
                assert np.allclose(x, y, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(x).max())), (cta, stash, r, field)


def test_simulate_cap_overflow_retry(eng):
    phase = cases.phase("si")
    gs, gt = _gtable_new_api(eng, phase, 5.0, True)
    wl = K.get_electron_wavelength(300)
    q = np.array([[1.0, 0, 0, 0]])
    spots = eng.simulate(gt, q, wl, 0.01, 0.01, "lorentzian", cap=32)
    assert spots.cap >= 70 and int(spots.count[0]) == 70


# --------------------------------------------------------------------------- K3
def _render_one(eng, xyz, inten, shape, **kw):
    import torch
    n = len(inten)
    cap = max(32, (n + 31) // 32 * 32)
    X = np.zeros((1, cap, 3))
    X[0, :n] = xyz
    I = np.zeros((1, cap))
    I[0, :n] = inten
    dev = eng.device()
    out = eng.render(torch.tensor([n], dtype=torch.int32, device=dev), torch.as_tensor(X, device=dev),
                     torch.as_tensor(I, device=dev), shape, **kw)
    return out[0].cpu().numpy()


@pytest.mark.parametrize("name", list(cases.DETECTOR_CASES))
@pytest.mark.parametrize("normalize", [True, False])
def test_render_fast_matches_reference_golden(eng, golden_dir, name, normalize):
    """detector.npz holds outputs of the reference's own integer-branch rasteriser."""
    shape, sigma, n, seed = cases.DETECTOR_CASES[name]
    xy, inten = cases.detector_spots(shape, n, seed)
    ref = np.load(golden_dir / "detector.npz")[f"{name}_int"]
    if normalize:
        ref = ref / ref.max()
    xyz = np.concatenate([xy, np.zeros((n, 1))], axis=1)
    got = _render_one(eng, xyz, inten, shape, sigma=sigma, calibration=1.0, center=(0.0, 0.0),
                      fast=True, normalize=normalize)
    assert np.abs(got - ref).max() <= IMG_ATOL * ref.max()


@pytest.mark.parametrize("name", list(cases.DETECTOR_CASES))
def test_render_slow_matches_reference_golden(eng, golden_dir, name):
    shape, sigma, n, seed = cases.DETECTOR_CASES[name]
    xy, inten = cases.detector_spots(shape, n, seed)
    ref = np.load(golden_dir / "detector.npz")[f"{name}_float"]
    xyz = np.concatenate([xy, np.zeros((n, 1))], axis=1)
    got = _render_one(eng, xyz, inten * 2000.0, shape, sigma=sigma, calibration=1.0, center=(0.0, 0.0),
                      fast=False, normalize=False, clip_threshold=1.0)
    assert np.abs(got - ref).max() <= IMG_ATOL * ref.max()


@pytest.mark.parametrize("shape,sigma,cal,angle,mirrored,fast", [
    ((256, 256), 10, 1 / 128, 0, False, True),
    ((256, 256), 1.4, 1 / 128, 30, True, True),
    ((144, 144), 10, 0.01, 0, False, True),
    ((100, 180), 3, 0.012, 77.5, False, True),
    ((256, 256), 4, 1 / 128, 12, True, False),
    ((144, 144), 10, 0.01, 0, False, False),
])
def test_render_matches_oracle_on_simulated_spots(eng, shape, sigma, cal, angle, mirrored, fast):
    phase = cases.phase("si")
    gs = K.GSet(phase.structure, 1.0, True)
    wl = K.get_electron_wavelength(200)
    for seed in range(3):
        G = oracle_G_from_active_quat(random_quats(1, seed)[0]) if seed else np.eye(3)
        r = K.simulate_rotation(phase.structure, gs, G, wl, 0.02)
        inten = r["intensity"] * (1.0 if fast else 50.0)
        center = (shape[1] // 2, shape[0] // 2) if fast else ((shape[1] - 1) / 2, (shape[0] - 1) / 2)
        ref = K.diffraction_pattern(r["xyz"], inten, shape, sigma=sigma, in_plane_angle=angle,
                                    calibration=cal, mirrored=mirrored, fast=fast, normalize=True)
        got = _render_one(eng, r["xyz"], inten, shape, sigma=sigma, calibration=cal, center=center,
                          in_plane_angle=angle, mirrored=mirrored, fast=fast, normalize=True)
        assert np.nanmax(np.abs(got - ref)) <= IMG_ATOL


def test_render_empty_is_zero(eng):
    got = _render_one(eng, np.zeros((0, 3)), np.zeros(0), (64, 64), sigma=3, calibration=0.01,
                      center=(32, 32), fast=True, normalize=True)
    assert np.all(got == 0)
    # all spots out of frame
    got = _render_one(eng, np.array([[5.0, 5.0, 0]]), np.array([1.0]), (64, 64), sigma=3, calibration=0.01,
                      center=(32, 32), fast=True, normalize=True)
    assert np.all(got == 0)


def test_render_graphite_golden(eng, golden_dir):
    """The reference's strongest pin (test_simulation_generator.py:283-343) through K2 + K3."""
    phase = cases.phase("graphite")
    gs, gt = _gtable_new_api(eng, phase, 1.6768, False)
    wl = K.get_electron_wavelength(200)
    G = K.bunge_matrix(*np.deg2rad([0, 90, 90]))
    from diffsims_b200.crystal import Rotation
    q = (~Rotation.from_euler([[0, 90, 90]], degrees=True)).data
    spots = eng.simulate(gt, q, wl, 0.1, 0.1, "lorentzian")
    assert int(spots.count[0]) == 104
    img = eng.render(spots.count, spots.xyz, spots.intensity, (128, 128), sigma=1.4, calibration=0.0262,
                     center=(64, 64))[0].cpu().numpy()
    old = np.load(golden_dir / "old_simulation.npz")["image"]
    assert np.abs(img - old).max() <= IMG_ATOL


# --------------------------------------------------------------------------- K3 schedule variants
@pytest.mark.parametrize("variant", ["umma", "umma_nowin", "rows", "pipe", "pipe_tma", "G8", "G4", "G2", "G1"])
@pytest.mark.parametrize("shape,sigma", [((256, 256), 10.0), ((144, 144), 3.0), ((90, 130), 2.0)])
def test_render_schedule_variants_agree_with_oracle(eng, opts, variant, shape, sigma):
    """The tcgen05 kernel, the warp-specialised pipelined kernel and every group size of the phase-synchronous
    kernel against the oracle (130 is not a multiple of 4: scalar-store instantiation)."""
    import torch
    if variant == "rows":     # the row-binned banded product (render_rows.cu)
        opts(render_group=-1, render_umma=1, render_rows=1)
    elif variant.startswith("umma"):
        opts(render_group=-1, render_umma=1, render_rows=0, render_umma_window=0 if variant == "umma_nowin" else -1)
    elif variant.startswith("pipe"):   # all-zero regions by st.global (default) or by TMA store (option; measured slower)
        opts(render_group=-1, render_pipe=1, render_umma=0, render_zero_tma=1 if variant == "pipe_tma" else -1)
    else:
        opts(render_group=int(variant[1:]), render_umma=0)
    phase = cases.phase("fe3c")
    gs = K.GSet(phase.structure, 1.2, True)
    wl = K.get_electron_wavelength(200)
    q = random_quats(40, 21)   # more templates than one CTA holds in flight: exercises the ticket hand-out
    cal = 1.2 / (min(shape) // 2)
    refs, X, I, cnt = [], np.zeros((len(q), 64, 3)), np.zeros((len(q), 64)), np.zeros(len(q), np.int32)
    for r in range(len(q)):
        ref = K.simulate_rotation(phase.structure, gs, oracle_G_from_active_quat(q[r]), wl, 0.01)
        n = len(ref["intensity"])
        assert n <= 64
        X[r, :n], I[r, :n], cnt[r] = ref["xyz"], ref["intensity"], n
        refs.append(K.diffraction_pattern(ref["xyz"], ref["intensity"], shape, sigma=sigma, calibration=cal))
    dev = eng.device()
    for normalize in (True, False):
        out = eng.render(torch.as_tensor(cnt, device=dev), torch.as_tensor(X, device=dev),
                         torch.as_tensor(I, device=dev), shape, sigma, cal, (shape[1] // 2, shape[0] // 2),
                         normalize=normalize).cpu().numpy()
        for r in range(len(q)):
            ref = refs[r] if normalize else K.diffraction_pattern(X[r, :cnt[r]], I[r, :cnt[r]], shape, sigma=sigma,
                                                                  calibration=cal, normalize=False)
            assert np.abs(out[r] - ref).max() <= IMG_ATOL * max(ref.max(), 1e-30), (variant, r)
    # twice on the same stream: the ticket words re-arm themselves
    again = eng.render(torch.as_tensor(cnt, device=dev), torch.as_tensor(X, device=dev),
                       torch.as_tensor(I, device=dev), shape, sigma, cal, (shape[1] // 2, shape[0] // 2),
                       normalize=False).cpu().numpy()
    np.testing.assert_array_equal(again, out)


def test_edge_sizes(eng):
    """Zero rotations, zero templates, an empty reciprocal set."""
    import torch
    phase = cases.phase("si")
    gs, gt = _gtable_new_api(eng, phase, 1.0, True)
    wl = K.get_electron_wavelength(200)
    spots = eng.simulate(gt, np.zeros((0, 4)), wl, 0.01, 0.01, "lorentzian")
    assert spots.n_rot == 0
    out = eng.render(spots.count, spots.xyz, spots.intensity, (64, 64), 3.0, 0.01, (32, 32))
    assert out.shape == (0, 64, 64)
    r, t, i = eng.polar_flatten(spots.count, spots.xyz, spots.intensity, 0)
    assert r.shape == (0, 0)
    # a reciprocal radius so small that only the direct beam is inside the sphere
    gs0, gt0 = _gtable_new_api(eng, phase, 0.05, True)
    assert gt0.n == 2
    s0 = eng.simulate(gt0, random_quats(3, 1), wl, 0.01, 0.01, "lorentzian")
    assert s0.count.tolist() == [2, 2, 2]
    gs1, gt1 = _gtable_new_api(eng, phase, 0.05, False)
    assert gt1.n == 0
    s1 = eng.simulate(gt1, random_quats(3, 1), wl, 0.01, 0.01, "lorentzian")
    assert s1.count.tolist() == [0, 0, 0]


def test_render_slow_path_with_many_spots(eng):
    """fast=False with a capacity above 32 (group size heuristics pick G = 2/4 for the fast path only; the
    sub-pixel path exists for G = 1 and 8) -- found by compute-sanitizer memcheck."""
    import torch
    rng = np.random.default_rng(7)
    n, cap, shape = 6, 96, (128, 128)
    X = np.zeros((n, cap, 3))
    X[..., :2] = rng.uniform(-0.9, 0.9, (n, cap, 2))
    I = rng.uniform(50, 500, (n, cap))
    cnt = np.array([96, 70, 33, 1, 0, 96], np.int32)
    dev = eng.device()
    out = eng.render(torch.as_tensor(cnt, device=dev), torch.as_tensor(X, device=dev), torch.as_tensor(I, device=dev),
                     shape, 2.5, 1 / 64, (63.5, 63.5), fast=False, normalize=False, clip_threshold=1.0).cpu().numpy()
    for r in range(n):
        ref = K.diffraction_pattern(X[r, :cnt[r]], I[r, :cnt[r]], shape, sigma=2.5, calibration=1 / 64, fast=False,
                                    normalize=False, clip_threshold=1.0) if cnt[r] else np.zeros(shape)
        assert np.abs(out[r] - ref).max() <= IMG_ATOL * max(ref.max(), 1.0)


@pytest.mark.parametrize("pipe", ["umma", "rows", "1", "0"])
@pytest.mark.parametrize("cap,sigma,normalize,shape", [
    (288, 10.0, True, (256, 256)), (288, 3.0, False, (256, 256)), (512, 6.0, True, (256, 256)),
    (1024, 2.0, True, (256, 256)), (288, 7.0, True, (96, 200)), (160, 5.0, True, (100, 150)),
    (96, 12.0, False, (40, 72))])
def test_render_fast_path_with_many_spots(eng, opts, pipe, cap, sigma, normalize, shape):
    """Dense patterns (hundreds of reflections per template) through all schedules: the tcgen05 kernel (capacities
    up to 1024, images up to 256 x 256), the warp-specialised kernel (capacities up to 512 while its slots fit in
    shared memory) and render_kernel for everything else.  The tensor-core paths use bf16 x 3 split products,
    incl. reflect images at the borders, partial tiles and rows that are not 16-byte multiples."""
    import torch
    if pipe in ("umma", "rows"):
        opts(render_umma=1, render_rows=int(pipe == "rows"))
    else:
        opts(render_umma=0, render_pipe=int(pipe))
    rng = np.random.default_rng(cap)
    n = 5
    X = np.zeros((n, cap, 3))
    X[..., 0] = rng.uniform(-1.05, 1.05, (n, cap)) * shape[1] / 256     # some outside the frame
    X[..., 1] = rng.uniform(-1.05, 1.05, (n, cap)) * shape[0] / 256
    I = rng.uniform(1, 500, (n, cap))
    cnt = np.array([cap, cap - 37, cap // 2 + 1, 1, 0], np.int32)
    dev = eng.device()
    centre = ((shape[1] - 1) / 2, (shape[0] - 1) / 2)
    out = eng.render(torch.as_tensor(cnt, device=dev), torch.as_tensor(X, device=dev), torch.as_tensor(I, device=dev),
                     shape, sigma, 1 / 128, centre, normalize=normalize).cpu().numpy()
    for r in range(n):
        if cnt[r] == 0:
            assert not out[r].any()
            continue
        ref = K.diffraction_pattern(X[r, :cnt[r]], I[r, :cnt[r]], shape, sigma=sigma, calibration=1 / 128,
                                    direct_beam_position=centre, normalize=normalize)
        assert np.abs(out[r] - ref).max() <= IMG_ATOL * max(ref.max(), 1.0)
        if normalize and ref.max() > 0:   # (the lone spot of template 3 can fall outside the frame: all zeros)
            assert out[r].max() == 1.0


def test_render_tensor_path_accuracy(eng, opts):
    """The bf16 split-product tensor-core paths (tcgen05 kernel; legacy mma.sync regions) against the float32 FMA
    path on the same dense templates: the difference stays below 3e-5 of the peak (budget: 1e-4), and the
    normalised maximum is exactly 1 in all of them."""
    import torch
    rng = np.random.default_rng(11)
    n, cap, shape = 4, 512, (256, 256)
    X = np.zeros((n, cap, 3))
    X[..., :2] = rng.uniform(-1.0, 1.0, (n, cap, 2))
    I = rng.uniform(1, 5000, (n, cap))
    cnt = torch.tensor([cap, 400, 300, 200], dtype=torch.int32, device=eng.device())
    args = (cnt, torch.as_tensor(X, device=eng.device()), torch.as_tensor(I, device=eng.device()), shape, 10.0,
            1 / 128, (127.5, 127.5))
    out = {}
    for name, kw in (("umma", dict(render_umma=1, render_rows=0, render_umma_team=4)),
                     ("umma2", dict(render_umma=1, render_rows=0, render_umma_team=2)), ("rows", dict(render_umma=1, render_rows=1)),
                     ("mma", dict(render_umma=0, render_mma=1)), ("fma", dict(render_umma=0, render_mma=0))):
        opts(**kw)
        out[name] = eng.render(*args).cpu().numpy()
    # two or four producer warps per operand stage: the same operands, the same products
    assert np.array_equal(out["umma"], out.pop("umma2"))
    for name in ("umma", "rows", "mma"):
        assert not np.array_equal(out[name], out["fma"])            # the tensor-core path did run
        assert np.abs(out[name] - out["fma"]).max() < 3e-5, name
    assert all(o[r].max() == 1.0 for o in out.values() for r in range(n))


@pytest.mark.parametrize("name,rr,model,min_int", [("si", 2.0, "lorentzian", 1e-20), ("al", 2.0, "linear", 1e-20),
                                                    ("fe3c", 1.5, "lorentzian", 1e-2), ("fe_bcc", 2.0, "atanc", 1e-4),
                                                    ("triclinic", 1.2, "binary", 0.2)])
def test_extinction_marking_is_exact(eng, opts, name, rr, model, min_int):
    """Rows that can never pass minimum_intensity are either dropped from the plan (compact=True, the default) or
    marked for K2's float32 cull (compact=False, ds_pack_gtable).  Both change nothing but the work: same
    reflections, same order, same numbers -- in the brute-force and the scan-line cull, for cuts from 1e-20
    (systematic absences only) up to 0.2 (most of the pattern)."""
    import diffsims_b200 as ds
    from diffsims_b200.utils import shape_factor_models as sfm
    phase = cases.phase(name)
    # (the reference accepts "binary" only as a callable)
    gen = ds.SimulationGenerator(200, shape_factor_model=sfm.binary if model == "binary" else model,
                                 minimum_intensity=min_int)
    assert gen._extinct_rel_cut(True) == pytest.approx(0.5 * min_int) and gen._extinct_rel_cut(False) == 0.0
    plan = gen._g_plan(phase, rr, True, {})
    rel = gen._extinct_rel_cut(True)
    q = random_quats(96, 4)
    q[0] = (1, 0, 0, 0)
    res = {}
    for lines in ("0", "1"):
        opts(sim_lines=int(lines))
        for mode in ("full", "marked", "compact"):
            gt = plan.run(0.0) if mode == "full" else plan.run(rel, compact=(mode == "compact"))
            n_dead = int(torch_isinf_count(gt.f32))
            assert (n_dead > 0) == (mode == "marked")
            assert (gt.n < plan.hkl.shape[0]) == (mode == "compact")
            res[lines, mode] = (gt, eng.simulate(gt, q, gen.wavelength, 0.02, 0.02, model, min_intensity=min_int,
                                                 want_exc=True))
    gt0, ref = res["0", "full"]
    assert int(ref.count.sum()) > 0
    for key, (gt, sp) in res.items():
        assert np.array_equal(sp.count.cpu().numpy(), ref.count.cpu().numpy()), key
        for r in range(len(q)):
            n = int(ref.count[r])
            assert np.array_equal(gt.hkl[sp.g_index[r, :n].cpu().numpy()],
                                  gt0.hkl[ref.g_index[r, :n].cpu().numpy()]), (key, r)
            for field in ("xyz", "intensity", "exc"):
                x, y = getattr(sp, field)[r, :n].cpu().numpy(), getattr(ref, field)[r, :n].cpu().numpy()
                assert np.allclose(x, y, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(y).max())), (key, r, field)


def torch_isinf_count(t):
    import torch
    return torch.isinf(t[:, 3]).sum().item()
