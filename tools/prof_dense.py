"""One dense K3 configuration for ncu (tcgen05 kernel forced):  python tools/prof_dense.py [fe3c|large|si] [n]"""
import sys
import torch
sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import _cabi, engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats

which = sys.argv[1] if len(sys.argv) > 1 else "fe3c"
cfg = {"fe3c": ("fe3c", 2.0, 0.05, 8192), "large": ("large", 2.5, 0.01, 2048), "si": ("si", 2.0, 0.05, 16384),
       "sparse": ("si", 1.0, 0.01, 32768)}[which]
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg[3]
dev = engine.device()
_cabi.set_option("render_umma", 1)
gen = ds.SimulationGenerator(200)
b = TemplateLibraryBuilder(gen, cases.phase(cfg[0]), reciprocal_radius=cfg[1], max_excitation_error=cfg[2], sigma=10.0,
                           calibration=cfg[1] / 128)
b.prepare()
q = torch.as_tensor(active_quaternions(random_quats(n, 0)), device=dev)
b.calibrate_cap(q)
sp = b.simulate(q)
img = torch.empty((n, 256, 256), dtype=torch.float32, device=dev)
for _ in range(3):
    b.render(sp, img)
torch.cuda.synchronize()
print("done", which, n, b.cap)
