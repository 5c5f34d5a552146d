"""Per-role cycle counters of the row-binned tcgen05 render kernel (CTA 0), from the DS_PROF build:
    python -c "from diffsims_b200 import build; build.build(defines=['DS_PROF'], out='tools/microbench/_libprof.so')"
    DIFFSIMS_B200_LIB=tools/microbench/_libprof.so python tools/prof_rows_roles.py
"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import _cabi, engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats

CONFIGS = [("C2 Si rr1 s.01", "si", 200, 1.0, 0.01, 10.0, 32768), ("C2 Si rr2 s.05", "si", 200, 2.0, 0.05, 10.0, 16384),
           ("C5 Fe3C rr2 s.05", "fe3c", 200, 2.0, 0.05, 10.0, 16384), ("C4 large rr2.5 s.01", "large", 200, 2.5, 0.01, 10.0, 8192)]
WAITS = {"front": ("slot_empty", "-", "-"), "mma": ("slot_full", "stage_full", "half_empty"),
         "epilogue": ("slot_full", "half_full", "max barrier"), "producer0": ("slot_full", "stage_empty", "-")}
dev = engine.device()
L = _cabi.lib()
_cabi.set_option("render_umma", 1)
_cabi.set_option("render_rows", 1)
for name, ph, kv, rr, s_max, sigma, n in CONFIGS:
    gen = ds.SimulationGenerator(kv)
    b = TemplateLibraryBuilder(gen, cases.phase(ph), reciprocal_radius=rr, max_excitation_error=s_max, sigma=sigma,
                               calibration=rr / 128)
    b.prepare()
    q = torch.as_tensor(active_quaternions(random_quats(n, 0)), device=dev)
    b.calibrate_cap(q)
    sp = b.simulate(q)
    img = torch.empty((n, 256, 256), dtype=torch.float32, device=dev)
    out = (ctypes.c_ulonglong * 16)()
    b.render(sp, img)
    b.render(sp, img)
    torch.cuda.synchronize()
    assert L.ds_debug_rows_prof(out) == 0
    v = np.array(list(out), dtype=np.float64).reshape(4, 4)
    per = n / 148.0
    print(f"{name:22s} templates/CTA ~{per:6.1f}")
    for i, r in enumerate(("front", "mma", "epilogue", "producer0")):
        print(f"    {r:10s} {v[i, 0] / per:8.0f} cyc/tmpl | waiting: " + ", ".join(
            f"{w} {v[i, 1 + j] / per:7.0f}" for j, w in enumerate(WAITS[r]) if w != "-"))
