"""Per-kernel timing over the BASELINE.json configs (device-resident inputs, CUDA events).
    python tools/bench_configs.py [n_rot]
"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats

n_rot = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
PEAK = 6553.6


def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


CONFIGS = [
    # name, phase, kV, rr, s_max, model, sigma, n_rot scale
    ("C2 Si rr1 s.01", "si", 200, 1.0, 0.01, "lorentzian", 10.0, 1.0),
    ("C2 Si rr1 s.05", "si", 200, 1.0, 0.05, "lorentzian", 10.0, 1.0),
    ("C2 Si rr2 s.01", "si", 200, 2.0, 0.01, "lorentzian", 10.0, 1.0),
    ("C2 Si rr2 s.05", "si", 200, 2.0, 0.05, "lorentzian", 10.0, 0.5),
    ("C2 Si rr2 s.05 sig1.4", "si", 200, 2.0, 0.05, "lorentzian", 1.4, 0.5),
    ("C3 Ti rr1 s.01", "ti", 300, 1.0, 0.01, "lorentzian", 10.0, 1.0),
    ("C5 Fe3C rr1 s.01", "fe3c", 200, 1.0, 0.01, "lorentzian", 10.0, 1.0),
    ("C5 Fe3C rr2 s.05", "fe3c", 200, 2.0, 0.05, "lorentzian", 10.0, 0.25),
    ("C4 large rr2.5 s.01", "large", 200, 2.5, 0.01, "lorentzian", 10.0, 1 / 64),
]
dev = engine.device()
for name, ph, kv, rr, s_max, model, sigma, scale in CONFIGS:
    phase = cases.phase(ph)
    n = max(64, int(n_rot * scale))
    gen = ds.SimulationGenerator(kv, shape_factor_model=model)
    b = TemplateLibraryBuilder(gen, phase, reciprocal_radius=rr, max_excitation_error=s_max, sigma=sigma,
                               calibration=rr / 128)
    b.prepare()
    t1 = timeit(lambda: b.prepare())
    q = torch.as_tensor(active_quaternions(random_quats(n, 0)), device=dev)
    b.calibrate_cap(q)
    sp = b.simulate(q)
    b.assert_no_overflow(sp)
    t2 = timeit(lambda: b.simulate(q))
    img = torch.empty((n, 256, 256), dtype=torch.float32, device=dev)
    t3 = timeit(lambda: b.render(sp, img))
    natoms = len(phase.structure)
    gbs = n * 262144 / t3 / 1e6
    print(f"{name:24s} n_g={b.plan.hkl.shape[0]:6d} live={b.gtable.n:6d} atoms={natoms:3d} n={n:6d} cap={b.cap:4d} spots/t={sp.count.float().mean().item():6.1f} | "
          f"K1 {t1*1e3:8.1f} us ({b.gtable.n*natoms/t1/1e6:7.2f} Gpair/s) | K2 {t2*1e3:8.1f} us ({n/t2/1e3:7.2f} Mrot/s) | "
          f"K3 {t3*1e3:8.1f} us ({n/t3/1e3:6.2f} Mtmpl/s, {gbs:5.0f} GB/s = {gbs/PEAK:5.1%})")
