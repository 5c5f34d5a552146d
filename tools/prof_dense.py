import sys, torch, numpy as np
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats
name, rr, s_max, n = sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
gen = ds.SimulationGenerator(200)
b = TemplateLibraryBuilder(gen, cases.phase(name), reciprocal_radius=rr, max_excitation_error=s_max, sigma=10.0, calibration=rr / 128)
b.prepare()
q = torch.as_tensor(active_quaternions(random_quats(n, 0)), device=engine.device())
b.calibrate_cap(q)
sp = b.simulate(q)
img = torch.empty((n, 256, 256), dtype=torch.float32, device=engine.device())
for _ in range(3):
    b.render(sp, img)
torch.cuda.synchronize()
