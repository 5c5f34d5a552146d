"""Randomised differential test: K1 + K2 + K3 against the oracle over seeded random configurations
(structure, voltage, reciprocal radius, excitation error, shape factor, precession, rotation, detector)."""
import numpy as np
import pytest

import diffsims_b200 as ds
from diffsims_b200 import engine
from diffsims_b200.library import active_quaternions
from diffsims_b200.utils import shape_factor_models as sfm
from oracle import kinematical as K
from tests.golden import cases
from tests.helpers import IMG_ATOL, compare_spots, random_quats

pytestmark = pytest.mark.gpu

MODELS = ["lorentzian", "linear", "sinc", "sin2c", "atanc", "binary"]


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration_matches_oracle(seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(1000 + seed)
    name = ["si", "al", "graphite", "ti", "fe_bcc", "fe_fcc", "fe3c", "triclinic"][seed % 8]
    phase = cases.phase(name)
    kv = float(rng.choice([80, 120, 200, 300]))
    rr = float(rng.uniform(0.6, 1.8))
    s_max = float(rng.choice([0.005, 0.01, 0.02, 0.05]))
    width = s_max if rng.random() < 0.5 else float(s_max * rng.uniform(0.5, 2.0))
    model = MODELS[int(rng.integers(len(MODELS)))]
    direct_beam = bool(rng.random() < 0.7)
    sp = ["lobato", "xtables"][int(rng.integers(2))]
    min_int = float(rng.choice([1e-20, 1e-6, 1e-3]))
    prec = float(rng.choice([0.0, 0.0, 0.3]))         # degrees; closed form when on
    kwargs = {"minima_number": int(rng.integers(3, 8))} if model in ("sinc", "sin2c", "atanc") and rng.random() < 0.5 else {}
    # "binary" is not a valid model STRING in the reference either: it is passed as the function
    gen = ds.SimulationGenerator(kv, scattering_params=sp, shape_factor_model=sfm.binary if model == "binary" else model,
                                 minimum_intensity=min_int,
                                 precession_angle=prec, approximate_precession=True, **kwargs)
    rot = ds.Rotation(random_quats(6, seed))
    sim = gen.calculate_diffraction2d(phase, rot, reciprocal_radius=rr, with_direct_beam=direct_beam,
                                      max_excitation_error=s_max, shape_factor_width=width,
                                      debye_waller_factors=cases.DW)
    gs = K.GSet(phase.structure, rr, direct_beam)
    hkl_ref = np.vstack([gs.hkl_int, [0, 0, 0]]) if direct_beam else gs.hkl_int
    key = {}
    for j, h in enumerate(hkl_ref):
        key.setdefault(tuple(h), []).append(j)
    shape = [(256, 256), (144, 144), (128, 200)][seed % 3]
    sigma = float(rng.choice([1.4, 4.0, 10.0]))
    cal = rr / (min(shape) // 2) * float(rng.uniform(0.8, 1.3))
    angle = float(rng.choice([0.0, rng.uniform(0, 360)]))
    mirrored = bool(rng.random() < 0.3)
    imgs = sim.get_diffraction_patterns(shape=shape, sigma=sigma, calibration=cal, in_plane_angle=angle,
                                        mirrored=mirrored).cpu().numpy()
    G = rot.to_matrix()
    F000sq = float(K.kinematical_intensities(phase.structure, np.zeros((1, 3)), np.zeros(1), cases.DW, sp)[0])
    for i, dv in enumerate(sim):
        ref = K.simulate_rotation(phase.structure, gs, G[i], gen.wavelength, s_max,
                                  K.SHAPE_FACTOR_MODELS[model], width, sp, cases.DW, min_int, prec, True, kwargs)
        seen, gidx = {}, []
        for h in dv.hkl.astype(int):       # the duplicated (000) maps to the two last table rows, in order
            t = tuple(h)
            gidx.append(key[t][seen.get(t, 0)])
            seen[t] = seen.get(t, 0) + 1
        got = dict(g_index=np.array(gidx, dtype=int), xyz=dv.data, intensity=dv.intensity,
                   excitation_error=np.zeros(dv.size))
        # reflections within 1e-6 of the cut may differ: identify them through the oracle's excitation errors
        near = {int(k) for k, s in zip(ref["g_index"], ref["excitation_error"])
                if prec == 0 and abs(abs(s) - s_max) < 1e-6}
        keep_ref = np.array([k not in near for k in ref["g_index"]], dtype=bool)
        keep_got = np.array([k not in near for k in got["g_index"]], dtype=bool)
        ref_f = {k: np.asarray(v)[keep_ref] for k, v in ref.items() if k in ("g_index", "xyz", "intensity", "excitation_error")}
        got_f = {k: np.asarray(v)[keep_got] for k, v in got.items()}
        # relative threshold: a reflection may sit exactly at max * min_int; compare above a small guard band
        big = ref_f["intensity"].max() if ref_f["intensity"].size else 0.0
        compare_spots(ref_f, got_f, s_max=-1.0, rr=rr, prec=True, noise_floor=max(1e-10, min_int * 1.001),
                      abs_floor=1e-12 * F000sq)
        if ref["intensity"].size and not near and ref["intensity"].max() > 1e-12 * F000sq:
            img = K.diffraction_pattern(ref["xyz"], ref["intensity"], shape, sigma=sigma, calibration=cal,
                                        in_plane_angle=angle, mirrored=mirrored)
            d = np.abs(imgs[i] - img)
            if np.nanmax(d) > IMG_ATOL:
                # the only accepted cause: a spot whose pixel coordinate sits within 1e-9 of a pixel boundary
                t = K.transformed_coordinates(ref["xyz"], angle, (shape[1] // 2, shape[0] // 2), mirrored, cal)
                frac = np.abs(t[:, :2] - np.rint(t[:, :2]))
                assert frac.min() < 1e-9, (seed, i, float(np.nanmax(d)))
