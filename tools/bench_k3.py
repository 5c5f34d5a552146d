"""K3-only timing on the headline workload with toggles: python tools/bench_k3.py [normalize 0/1] [n]"""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats
norm = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
gen = ds.SimulationGenerator(200)
b = TemplateLibraryBuilder(gen, cases.phase("si"), reciprocal_radius=1.0, max_excitation_error=0.01, sigma=10.0,
                           calibration=1 / 128, normalize=norm)
b.prepare()
q = torch.as_tensor(active_quaternions(random_quats(n, 0)), device=engine.device())
b.calibrate_cap(q)
sp = b.simulate(q)
img = torch.empty((n, 256, 256), dtype=torch.float32, device=engine.device())
ts = []
for i in range(8):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); b.render(sp, img); e.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(e))
t = float(np.median(ts[2:]))
print(f"normalize={norm} K3 {t*1e3:.1f} us  {n*262144/t/1e6:.0f} GB/s  {n*262144/t/1e6/6553.6:.1%}")
