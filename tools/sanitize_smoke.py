"""One small launch of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import _cabi, engine
from diffsims_b200.library import pack_csr
from diffsims_b200.generators.rotation_list_generators import beam_directions_device
from tests.golden import cases
from tests.helpers import random_quats

gen = ds.SimulationGenerator(200)
for name, rr in (("si", 1.0), ("large", 1.2)):          # resident table and streaming tiles
    gt = gen._g_table(cases.phase(name), rr, True, cases.DW)
    q = random_quats(24, 0)
    for model, prec in (("lorentzian", 0.0), ("sinc", 0.0), ("lorentzian_precession", 0.0087), ("linear", 0.0087)):
        sp = engine.simulate(gt, q, gen.wavelength, 0.01, 0.01, model, precession_rad=prec, want_exc=True)
    for variant in ("umma", "pipe", "pipe_tma", "8", "2"):
        _cabi.set_option("render_group", int(variant) if variant in ("8", "2") else -1)
        _cabi.set_option("render_umma", 1 if variant == "umma" else 0)
        _cabi.set_option("render_zero_tma", 1 if variant == "pipe_tma" else -1)
        for fast in (True, False):
            for shape in ((256, 256), (70, 90)):
                engine.render(sp.count, sp.xyz, sp.intensity, shape, 6.0, rr / 64, (shape[1] // 2, shape[0] // 2), fast=fast)
    for k in ("render_group", "render_umma", "render_zero_tma"):
        _cabi.set_option(k, -1)
    pack_csr(sp)
    engine.render(sp.count, sp.xyz, sp.intensity, (24, 24), 10.0, rr / 12, (12, 12))     # kernel wider than the image
    engine.polar_flatten(sp.count, sp.xyz, sp.intensity, int(sp.count.max()), np.linspace(0, 1, 20), np.linspace(-3.2, 3.2, 30))
# dense patterns: tensor-core regions (hit lists, reflect images, lane-pair exchange), both schedules, partial regions;
# scan-line cull of K2; the bare rasteriser; mesh vertices through ds_beam_points
rng = np.random.default_rng(3)
for cap, shape in ((288, (256, 256)), (160, (100, 152)), (1024, (96, 200))):
    n = 3
    X = np.zeros((n, cap, 3))
    X[..., :2] = rng.uniform(-1.05, 1.05, (n, cap, 2)) * (shape[1] / 256, shape[0] / 256)
    I = rng.uniform(1, 500, (n, cap))
    cnt = torch.tensor([cap, cap // 2, 17], dtype=torch.int32, device=engine.device())
    # tcgen05 kernels (per-reflection product with 2- and 4-warp producer teams, row-binned banded product; + their
    # prepare passes), pipelined, phase-synchronous
    for umma, pipe, rows, team in ((1, -1, 0, 2), (1, -1, 0, 4), (1, -1, 1, -1), (0, 1, -1, -1), (0, 0, -1, -1)):
        _cabi.set_option("render_umma", umma)
        _cabi.set_option("render_pipe", pipe)
        _cabi.set_option("render_rows", rows)
        _cabi.set_option("render_umma_team", team)
        for normalize in (True, False):
            engine.render(cnt, torch.as_tensor(X, device=engine.device()), torch.as_tensor(I, device=engine.device()),
                          shape, 7.0, 1 / 128, ((shape[1] - 1) / 2, (shape[0] - 1) / 2), normalize=normalize)
for k in ("render_umma", "render_pipe", "render_rows", "render_umma_team"):
    _cabi.set_option(k, -1)
_cabi.set_option("sim_lines", 1)
gt = gen._g_table(cases.phase("si"), 2.0, True, cases.DW)
engine.simulate(gt, random_quats(40, 1), gen.wavelength, 0.01, 0.01, "lorentzian")
engine.simulate(gt, random_quats(40, 1), gen.wavelength, 0.01, 0.01, "lorentzian_precession", precession_rad=0.0087)
# extinct rows marked in the packed table (compact=False), culled by the plain and by the scan-line loop
plan = gen._g_plan(cases.phase("si"), 2.0, True, {})
for lines in (0, 1):
    _cabi.set_option("sim_lines", lines)
    engine.simulate(plan.run(0.5e-20, compact=False), random_quats(40, 2), gen.wavelength, 0.01, 0.01, "lorentzian")
engine.simulate(plan.run(0.5e-20), random_quats(40, 2), gen.wavelength, 0.01, 0.01, "lorentzian")
_cabi.set_option("sim_lines", -1)
# few rotations over a large table (one CTA per rotation: default stash, a small pool, the fall-back, brute-force and
# scan-line scans; 2 warps per CTA),
# the factorised structure factors (>= 4096 rows, >= 32 atoms: box kernel), the SO(3) grids
plan = gen._g_plan(cases.phase("large"), 1.6, True, cases.DW)
gt_large = plan.run(0.0)
for cta, stash, lines in ((-1, -1, -1), (1, 512, 0), (1, 512, 1), (1, 64, 0), (1, -1, 1), (0, -1, -1)):
    _cabi.set_option("sim_cta", cta)
    _cabi.set_option("sim_stash", stash)
    _cabi.set_option("sim_lines", lines)
    engine.simulate(gt_large, random_quats(40, 3), gen.wavelength, 0.01, 0.01, "lorentzian")
    engine.simulate(gt_large, random_quats(9, 4), gen.wavelength, 0.01, 0.01, "sin2c", precession_rad=0.005)
for k in ("sim_cta", "sim_stash", "sim_lines"):
    _cabi.set_option(k, -1)
from diffsims_b200.generators.rotation_list_generators import fundamental_zone_device, local_grid_device
fundamental_zone_device(12, point_group="m-3m")
local_grid_device(10, center=(10, 20, 30), grid_width=30)
from diffsims_b200.pattern.detector_functions import get_pattern_from_pixel_coordinates_and_intensities
get_pattern_from_pixel_coordinates_and_intensities(rng.uniform(-8, 98, (40, 2)), rng.uniform(20, 900, 40), (70, 90), 2.5)
beam_directions_device("cubic", 6.0, mesh="icosahedral")
e, qd = beam_directions_device("hexagonal", 3.0)
torch.cuda.synchronize()
print("sanitize smoke ok", e.shape)
