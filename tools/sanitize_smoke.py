"""One small launch of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import engine
from diffsims_b200.generators.rotation_list_generators import beam_directions_device
from tests.golden import cases
from tests.helpers import random_quats

gen = ds.SimulationGenerator(200)
for name, rr in (("si", 1.0), ("large", 1.2)):          # resident table and streaming tiles
    gt = gen._g_table(cases.phase(name), rr, True, cases.DW)
    q = random_quats(24, 0)
    for model, prec in (("lorentzian", 0.0), ("sinc", 0.0), ("lorentzian_precession", 0.0087), ("linear", 0.0087)):
        sp = engine.simulate(gt, q, gen.wavelength, 0.01, 0.01, model, precession_rad=prec, want_exc=True)
    for variant in ("pipe", "8", "2"):
        os.environ.pop("DS_RENDER_GROUP", None)
        if variant != "pipe":
            os.environ["DS_RENDER_GROUP"] = variant
        for fast in (True, False):
            for shape in ((256, 256), (70, 90)):
                engine.render(sp.count, sp.xyz, sp.intensity, shape, 6.0, rr / 64, (shape[1] // 2, shape[0] // 2), fast=fast)
    engine.render(sp.count, sp.xyz, sp.intensity, (24, 24), 10.0, rr / 12, (12, 12))     # kernel wider than the image
    engine.polar_flatten(sp.count, sp.xyz, sp.intensity, int(sp.count.max()), np.linspace(0, 1, 20), np.linspace(-3.2, 3.2, 30))
e, qd = beam_directions_device("hexagonal", 3.0)
torch.cuda.synchronize()
print("sanitize smoke ok", e.shape)
