"""K3 variants over the BASELINE.json configs (device-resident spot tables, CUDA events).
    python tools/bench_k3_variants.py [n_rot]
For each configuration: the two tcgen05 kernels (per-reflection rank-S product, row-binned banded product) against the float32 /
mma.sync kernels of round 1, as a fraction of the measured copy peak (MEASURED_PEAKS.json hbm_gbs)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, ".")
import diffsims_b200 as ds
from diffsims_b200 import _cabi, engine
from diffsims_b200.library import TemplateLibraryBuilder, active_quaternions
from tests.golden import cases
from tests.helpers import random_quats

import os
n_rot = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
NORMALIZE = os.environ.get("NORMALIZE", "1") != "0"
peaks = Path("MEASURED_PEAKS.json")
PEAK = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0


def timeit(f, n=5):
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


CONFIGS = [
    # name, phase, kV, rr, s_max, sigma, n_rot scale
    ("C2 Si rr1 s.01", "si", 200, 1.0, 0.01, 10.0, 1.0),
    ("C2 Si rr1 s.05", "si", 200, 1.0, 0.05, 10.0, 1.0),
    ("C2 Si rr2 s.01", "si", 200, 2.0, 0.01, 10.0, 1.0),
    ("C2 Si rr2 s.05", "si", 200, 2.0, 0.05, 10.0, 0.5),
    ("C2 Si rr2 s.05 sig1.4", "si", 200, 2.0, 0.05, 1.4, 0.5),
    ("C3 Ti rr1 s.01", "ti", 300, 1.0, 0.01, 10.0, 1.0),
    ("C5 Fe3C rr1 s.01", "fe3c", 200, 1.0, 0.01, 10.0, 1.0),
    ("C5 Fe3C rr2 s.05", "fe3c", 200, 2.0, 0.05, 10.0, 0.5),
    ("C4 large rr2.5 s.01", "large", 200, 2.5, 0.01, 10.0, 1 / 8),
]
VARIANTS = [("tcgen05 per-spot", dict(render_umma=1, render_rows=0)),
            ("tcgen05 rows", dict(render_umma=1, render_rows=1)),
            ("float32 kernels", dict(render_umma=0, render_rows=-1))]
dev = engine.device()
only = sys.argv[2] if len(sys.argv) > 2 else None
for name, ph, kv, rr, s_max, sigma, scale in CONFIGS:
    if only and only not in name:
        continue
    phase = cases.phase(ph)
    n = max(64, int(n_rot * scale))
    gen = ds.SimulationGenerator(kv)
    b = TemplateLibraryBuilder(gen, phase, reciprocal_radius=rr, max_excitation_error=s_max, sigma=sigma,
                               calibration=rr / 128, normalize=NORMALIZE)
    b.prepare()
    q = torch.as_tensor(active_quaternions(random_quats(n, 0)), device=dev)
    b.calibrate_cap(q)
    sp = b.simulate(q)
    b.assert_no_overflow(sp)
    img = torch.empty((n, 256, 256), dtype=torch.float32, device=dev)
    line = f"{name:24s} n={n:6d} cap={b.cap:4d} spots/t={sp.count.float().mean().item():6.1f} |"
    ref = None
    for vname, kw in VARIANTS:
        for k, v in kw.items():
            _cabi.set_option(k, v)
        t3 = timeit(lambda: b.render(sp, img))
        gbs = n * 262144 / t3 / 1e6
        line += f" {vname}: {t3 * 1e3:8.1f} us {gbs / PEAK:5.1%} |"
        got = img[: min(n, 64)].clone()
        if ref is None:
            ref = got
        else:
            line += f" dmax {float((got - ref).abs().max()):.1e} |"
    for k in ("render_umma", "render_rows"):
        _cabi.set_option(k, -1)
    print(line, flush=True)
