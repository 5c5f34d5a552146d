"""K3-only timing on any configuration: python tools/bench_k3_cfg.py phase rr s_max n [normalize 0/1] [sigma]
(environment toggles DS_RENDER_PIPE / DS_RENDER_GROUP / DS_RENDER_FRONTS / DS_RENDER_PIPE_MAXCAP apply)."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats
name, rr, s_max, n = sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
norm = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
sigma = float(sys.argv[6]) if len(sys.argv) > 6 else 10.0
gen = ds.SimulationGenerator(200)
b = TemplateLibraryBuilder(gen, cases.phase(name), reciprocal_radius=rr, max_excitation_error=s_max, sigma=sigma,
                           calibration=rr / 128, normalize=norm)
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
print(f"{name} rr={rr} s={s_max} n={n} cap={sp.cap} spots/t={float(sp.count.float().mean()):.1f} norm={norm} "
      f"K3 {t*1e3:.1f} us  {n*262144/t/1e6:.0f} GB/s  {n*262144/t/1e6/6553.6:.1%}")
