"""Scratch timing of K1/K2/K3 on the BASELINE config-2 workload (device-resident inputs)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from diffsims_b200 import engine
from diffsims_b200.crystallography import g_set_from_min_dspacing
from tests.golden import cases
from tests.helpers import random_quats

n_rot = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
s_max = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 10.0
phase = cases.phase("si")
lat = phase.structure.lattice
hkl = g_set_from_min_dspacing(lat, 1 / rr, True)
hkl = np.vstack([hkl, [0, 0, 0]])
xyz = hkl @ lat.recbase.T
dev = engine.device()

def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

t = timeit(lambda: engine.make_gtable(phase.structure, hkl, xyz, None, "lobato"))
print(f"K1+pack n_g={len(hkl)}: {t} ms")
gt = engine.make_gtable(phase.structure, hkl, xyz, None, "lobato")
q = torch.as_tensor(random_quats(n_rot, 0), device=dev)
wl = 0.025079340436272274
sp = engine.simulate(gt, q, wl, s_max, s_max, "lorentzian")
cap = sp.cap
print("cap", cap, "mean count", sp.count.float().mean().item(), "max", sp.count.max().item())
t = timeit(lambda: engine.simulate(gt, q, wl, s_max, s_max, "lorentzian", cap=cap, check_overflow=False))
print(f"K2 {n_rot} rot: {t} ms -> {n_rot / t[0] * 1e3 / 1e6:.2f} M rot/s")
out = torch.empty((n_rot, 256, 256), dtype=torch.float32, device=dev)
for norm in (True, False):
    t = timeit(lambda: engine.render(sp.count, sp.xyz, sp.intensity, (256, 256), sigma, rr / 128, (128, 128), normalize=norm, out=out))
    gbs = n_rot * 256 * 256 * 4 / t[0] / 1e6
    print(f"K3 normalize={norm} sigma={sigma}: {t} ms -> {n_rot / t[0] * 1e3 / 1e6:.2f} M tmpl/s, {gbs:.0f} GB/s ({gbs / 6553.6:.2%} of measured HBM)")
