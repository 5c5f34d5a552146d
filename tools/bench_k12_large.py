"""K1 / K2 on the large-cell config (BASELINE configs[3]): factorised vs direct structure factors, K2 per-rotation cost
for few and many rotations (one warp per rotation with 8 / 2 warps per CTA, one CTA per rotation).   python tools/bench_k12_large.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import _cabi, engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats


def timeit(f, n=7):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


dev = engine.device()
gen = ds.SimulationGenerator(200)
b = TemplateLibraryBuilder(gen, cases.phase("large"), reciprocal_radius=2.5, max_excitation_error=0.01, sigma=10.0, calibration=2.5 / 128)
b.prepare()
plan = b.plan
pairs = b.gtable.n * 500
t_fact = timeit(lambda: plan._run(0.0))
I_fact = plan._run(0.0).I0.clone()
saved = plan.sf_scratch
plan.sf_scratch = None
t_dir = timeit(lambda: plan._run(0.0))
I_dir = plan._run(0.0).I0.clone()
plan.sf_scratch = saved
rel = float(((I_fact - I_dir).abs() / I_dir.abs().clamp(min=1e-300)).max())
print(f"K1 large cell ({b.gtable.n} g x 500 atoms, incl. table packing): factorised {t_fact * 1e3:7.1f} us ({pairs / t_fact / 1e6:6.1f} Gpair/s) | "
      f"direct {t_dir * 1e3:7.1f} us ({pairs / t_dir / 1e6:6.1f} Gpair/s) | max rel diff of |F|^2 {rel:.2e}")

q_all = torch.as_tensor(active_quaternions(random_quats(16384, 0)), device=dev)
b.calibrate_cap(q_all[:2048])
VARIANTS = [("auto", dict(sim_cta=-1, sim_split=-1, sim_lines=-1)), ("warp/rot x8", dict(sim_cta=0, sim_split=8, sim_lines=0)),
            ("warp/rot x2", dict(sim_cta=0, sim_split=2, sim_lines=0)), ("CTA/rot", dict(sim_cta=1, sim_split=-1, sim_lines=0)),
            ("CTA/rot lines", dict(sim_cta=1, sim_split=-1, sim_lines=1))]
for n in (128, 512, 1024, 2048, 16384):
    q = q_all[:n].contiguous()
    line = f"K2 large cell n_rot={n:6d}:"
    for name, kw in VARIANTS:
        for k, v in kw.items():
            _cabi.set_option(k, v)
        t = timeit(lambda: b.simulate(q), n=5)
        line += f"  {name}: {t * 1e3:8.1f} us = {t * 1e6 / n:6.1f} ns/rot"
    for k in ("sim_cta", "sim_split", "sim_lines"):
        _cabi.set_option(k, -1)
    print(line, flush=True)
