#!/usr/bin/env python
"""Benchmark of the kinematical template-simulation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2_dense|c3|c4|c5] [--batch B] [--impl reference]

A STEP builds one template library of B orientations per GPU (default: the BASELINE configs[1] workload "c2": Si
diamond cubic, 200 kV, reciprocal_radius 1.0, max_excitation_error 0.01, lorentzian, lobato, direct beam, 256 x 256 px
float32 templates, sigma 10, calibration rr/128, normalised; B = 262 144):
    K1 structure factors + table packing -> K2 simulate (all rotations) -> K3 rasterise (all templates).
`value` is templates/s over all GPUs with the rotation list already in HBM; `e2e` is the same library built through
the public host-buffer call (TemplateLibraryBuilder.run_host): pinned host quaternions in, float32 images out into a
pinned host ring, every copy inside the timed region, next to the measured device->host roof of the box
(`d2h_roof`) and to `e2e_spots` -- host rotations in, packed spot lists (what calculate_diffraction2d returns) on the
host out.  `e2e_sharded` builds BASELINE configs[4] (Fe bcc + Fe fcc + Fe3C, 131 072 orientations per GPU) with
ShardedLibraryBuilder: slices balanced by cost, no data-path collective, one gather of the packed spot lists at the end
(its share is reported).  `extra` carries, per BASELINE config, the kernel times, K3's fraction of the HBM roof, K2's
rate and the spot-lists-only throughput.  Rotation lists shard across ranks (weak scaling: B per GPU).

`--impl reference` times the float64 CPU oracle (the port of the reference's numpy path; the reference itself
cannot be imported in this image) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "templates/sec (Si, 256x256 px)"
UNIT = "templates/s"
WORKLOAD = dict(
    workload="Si Fd-3m a=5.431 8 atoms, 200 kV, reciprocal_radius=1.0, max_excitation_error=0.01, "
             "lorentzian, lobato, direct beam, 256x256 float32 templates, sigma=10, calibration=1/128, "
             "normalised; uniform random orientations (BASELINE configs[1], synthetic rotation grid)",
    kv=200, rr=1.0, s_max=0.01, shape=(256, 256), sigma=10.0, calibration=1.0 / 128)

# The BASELINE.json configs as bench workloads.  phase: tests/golden/cases.py name; batch:
# orientations per GPU per step (sized so that a step is >= ~10 ms); c5 is a list of phases.
WORKLOADS = {
    "c2": dict(phases=["si"], kv=200, rr=1.0, s_max=0.01, sigma=10.0, batch=262144, text=WORKLOAD["workload"]),
    "c2_dense": dict(phases=["si"], kv=200, rr=2.0, s_max=0.05, sigma=10.0, batch=65536,
                     text="Si, 200 kV, reciprocal_radius=2.0, max_excitation_error=0.05 (39 reflections per template), "
                          "256x256 float32, sigma=10, normalised (BASELINE configs[1], dense variant)"),
    "c3": dict(phases=["ti"], kv=300, rr=1.0, s_max=0.01, sigma=10.0, batch=131072,
               text="hexagonal Ti P6_3/mmc, 300 kV, reciprocal_radius=1.0, max_excitation_error=0.01, lorentzian, "
                    "256x256 float32, sigma=10, normalised (BASELINE configs[2])"),
    "c4": dict(phases=["large"], kv=200, rr=2.5, s_max=0.01, sigma=10.0, batch=8192,
               text="synthetic cubic cell a=12 A with 500 atoms, reciprocal_radius=2.5 (113 082 g vectors, 680 reflections "
                    "per template), 256x256 float32, sigma=10, normalised (BASELINE configs[3])"),
    "c5": dict(phases=["fe_bcc", "fe_fcc", "fe3c"], kv=200, rr=1.0, s_max=0.01, sigma=10.0, batch=131072,
               text="multi-phase library Fe bcc + Fe fcc + Fe3C, equal orientation counts, 200 kV, reciprocal_radius=1.0, "
                    "max_excitation_error=0.01, 256x256 float32, sigma=10, normalised, sharded by cost with "
                    "ShardedLibraryBuilder (BASELINE configs[4])"),
}


def bench_phase(name):
    if name == "si":
        return si_phase()
    from tests.golden import cases
    return cases.phase(name)


def si_phase():
    from diffsims_b200.crystal import Atom, Lattice, Phase, Structure
    a = 5.431
    latt = Lattice(a, a, a, 90, 90, 90)
    atoms = []
    for c in [[0, 0, 0], [0.5, 0, 0.5], [0, 0.5, 0.5], [0.5, 0.5, 0]]:
        atoms.append(Atom("Si", c))
        atoms.append(Atom("Si", [c[0] + 0.25, c[1] + 0.25, c[2] + 0.25]))
    return Phase("Si", space_group=227, structure=Structure(atoms, latt))


def random_quats(n, seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    return q


# ------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(wname="c2"):
    """One (phase, g set, wavelength) entry per phase of the workload; c5 cycles through its three phases."""
    from oracle import kinematical as K
    w = WORKLOADS[wname]
    _CPU.clear()
    _CPU["K"] = K
    _CPU["w"] = w
    _CPU["wl"] = K.get_electron_wavelength(w["kv"])
    _CPU["phases"] = []
    for name in w["phases"]:
        phase = bench_phase(name)
        _CPU["phases"].append((phase, K.GSet(phase.structure, w["rr"], True)))


def _cpu_templates(quats):
    """Spot list + rendered template for each quaternion with the float64 oracle; returns a checksum."""
    if not _CPU:
        _cpu_init()
    K, w = _CPU["K"], _CPU["w"]
    acc = 0.0
    for i, q in enumerate(quats):
        phase, gs = _CPU["phases"][i % len(_CPU["phases"])]
        G = K.quat_to_matrix(q)  # passive matrix: rotated = g @ G
        r = K.simulate_rotation(phase.structure, gs, G, _CPU["wl"], w["s_max"])
        img = K.diffraction_pattern(r["xyz"], r["intensity"], WORKLOAD["shape"], sigma=w["sigma"],
                                    calibration=WORKLOAD["calibration"])
        acc += float(img[128, 128])
    return acc


def cpu_baseline_single_core(seconds=12.0):
    _cpu_init()
    q = random_quats(4096, 123)
    _cpu_templates(q[:4])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds and n < len(q):
        _cpu_templates(q[n:n + 16])
        n += 16
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit=UNIT, cores=1, kind="port",
                sample=f"{n} templates of the same workload (spot list + 256x256 sigma=10 render), "
                       f"float64 numpy oracle, 1 process, {dt:.1f} s")


def run_reference_arm(args):
    """`--impl reference`: the CPU oracle on every host core; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    wname = args.workload
    per_worker = {"c4": 2, "c5": 12, "c2_dense": 12}.get(wname, 24)  # a step stays a few seconds of CPU work
    sample = cores * per_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(wname,)) as pool:
        def step(seed):
            q = random_quats(sample, seed)
            chunks = [q[i * per_worker:(i + 1) * per_worker] for i in range(cores)]
            t0 = time.perf_counter()
            pool.map(_cpu_templates, chunks, chunksize=1)
            return time.perf_counter() - t0
        for w in range(args.warmup):
            step(1000 + w)
        times = [step(w) for w in range(args.steps)]
    total = sum(times)
    value = sample * args.steps / total
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=WORKLOADS[wname]["text"], workload_key=wname, templates_per_step=sample,
                            note="reference cannot be imported in this image (orix/diffpy absent); this is the "
                                 "float64 numpy port pinned by the reference's fixtures"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{sample} templates per step ({per_worker} per process x {cores} processes)"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks / throttle reasons sampled while the steps run (NVML every 5 ms, else nvidia-smi every 20 ms); the summary uses the samples taken inside the timed
    region (``mark_start`` .. ``mark_end``) and falls back to the whole window when the region is shorter
    than a sampling period."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.stop, self.index = [], threading.Event(), index
        self.t0 = self.t1 = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self._run_nvml():
            return
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                     "--format=csv,noheader,nounits", "-lms", "20"],
                                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        try:
            for line in proc.stdout:
                self.samples.append((time.perf_counter(), [f.strip() for f in line.split(",")]))
                if self.stop.is_set():
                    break
        finally:
            proc.kill()

    def _run_nvml(self):
        """The same fields through NVML in this thread (5 ms period; an nvidia-smi process answers too slowly on
        an 8-GPU box to land inside a 0.1 s timed region).  False if NVML is unavailable."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            reasons_of = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons",
                                 getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = (0x8, 0x40, 0x20, 0x4)   # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = int(reasons_of(h)) if reasons_of else 0
                self.samples.append((time.perf_counter(), [str(sm), str(mx)] +
                                     ["Active" if mask & b else "Not Active" for b in bits]))
            except Exception:
                pass
            time.sleep(0.005)
        return True

    def __enter__(self):
        self.thread.start()
        time.sleep(0.15)  # let nvidia-smi come up before the timed region starts
        return self

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        time.sleep(0.05)
        self.stop.set()
        self.thread.join(timeout=2)

    def summary(self):
        inside = [s for t, s in self.samples if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        use = inside if inside else [s for _, s in self.samples]
        num = lambda x: x.replace(".", "").isdigit()
        sm = [float(s[0]) for s in use if s and num(s[0])]
        mx = [float(s[1]) for s in use if len(s) > 1 and num(s[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i] == "Active" for s in use)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm), samples_in_timed_region=len(inside))


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (first-touch then places the pinned
    host buffers of the end-to-end path next to the GPU's PCIe root).  Best effort; returns the node or None."""
    try:
        out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]   # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = []
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time_ms(f, reps=5):
    """Median CUDA-event time of f() on the current stream (after one untimed call)."""
    import torch
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = _events()
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def make_builder(w, phase_name, dev):
    from diffsims_b200 import SimulationGenerator
    from diffsims_b200.library import TemplateLibraryBuilder
    gen = SimulationGenerator(w["kv"])
    return TemplateLibraryBuilder(gen, bench_phase(phase_name), reciprocal_radius=w["rr"], max_excitation_error=w["s_max"],
                                  shape=WORKLOAD["shape"], sigma=w["sigma"], calibration=w["rr"] / 128)


def config_probe(name, dev, peak, n=None):
    """Per-kernel figures of one BASELINE config on this GPU (device-resident, CUDA events): the `extra` block."""
    import torch
    from diffsims_b200 import engine
    from diffsims_b200.library import active_quaternions
    w = WORKLOADS[name]
    out = []
    for ph in w["phases"]:
        n_t = n or max(1024, w["batch"] // (4 if name != "c4" else 1))
        b = make_builder(w, ph, dev)
        b.prepare()
        q = torch.as_tensor(active_quaternions(random_quats(n_t, 7)), device=dev)
        b.calibrate_cap(q)
        sp = b.simulate(q)
        b.assert_no_overflow(sp)
        H, W = b.shape
        img = torch.empty((n_t, H, W), dtype=torch.float32, device=dev)
        k1 = _time_ms(lambda: b.prepare())
        k2 = _time_ms(lambda: b.simulate(q))
        k3 = _time_ms(lambda: b.render(sp, img))
        gbs = n_t * H * W * 4 / (k3 * 1e-3) / 1e9
        # spot lists only: K1 + K2 with the rotation list in HBM (what calculate_diffraction2d computes)
        spots_ms = _time_ms(lambda: (b.prepare(), b.simulate(q)))
        n_g = int(b.plan.hkl.shape[0])
        out.append(dict(config=name, phase=ph, templates=n_t, n_g=n_g, n_g_live=int(b.gtable.n),
                        atoms=len(b.phase.structure), mean_spots=float(sp.count.float().mean().item()), spot_capacity=int(b.cap),
                        k1_us=k1 * 1e3, k2_us=k2 * 1e3, k3_us=k3 * 1e3,
                        k1_gpairs_per_s=b.gtable.n * len(b.phase.structure) / (k1 * 1e-3) / 1e9,
                        k2_mrot_per_s=n_t / (k2 * 1e-3) / 1e6, k2_g_rows_per_s=n_t * b.gtable.n / (k2 * 1e-3),
                        k3_gbs=gbs, k3_frac=gbs / peak, k3_kernels=engine.render_launch_count(b.cap, b.shape, b.sigma, True, b.mean_spots),
                        templates_per_s=n_t / ((k1 + k2 + k3) * 1e-3), spot_lists_only_templates_per_s=n_t / (spots_ms * 1e-3)))
        if b.gtable.n >= 4096:
            # K2 over a large table: a launch of few rotations (one CTA per rotation, the table split across its warps)
            # against the warp-per-rotation kernel at scale
            q_big = torch.as_tensor(active_quaternions(random_quats(16384, 9)), device=dev)
            few = _time_ms(lambda: b.simulate(q_big[:512]))
            many = _time_ms(lambda: b.simulate(q_big))
            out[-1]["k2_large_table"] = dict(ns_per_rotation_512=few * 1e6 / 512, ns_per_rotation_16384=many * 1e6 / 16384,
                                             ratio=(few / 512) / (many / 16384))
            del q_big
        del img, sp, q
        torch.cuda.empty_cache()
    return out


def d2h_roof(dev, world, barrier, seconds=1.0, slot_bytes=1 << 30):
    """What the box gives for device->host copies: every rank copies a device buffer into pinned host slots at the
    same time, no kernels.  Returns GB/s of this rank (the caller sums over ranks)."""
    import torch
    src = torch.empty(slot_bytes, dtype=torch.uint8, device=dev)
    dst = [torch.empty(slot_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for d in dst:
        d.copy_(src, non_blocking=True)
    barrier()
    a, b = _events()
    n = 0
    t0 = time.perf_counter()
    a.record()
    while time.perf_counter() - t0 < seconds:
        for d in dst:
            d.copy_(src, non_blocking=True)
            n += 1
        torch.cuda.current_stream().synchronize()
    b.record()
    torch.cuda.synchronize()
    gbs = n * slot_bytes / (a.elapsed_time(b) * 1e-3) / 1e9
    del src, dst
    return gbs


def run_c5(args, rank, world, local_rank, dev, barrier, max_over_ranks, peak):
    """`--workload c5`: a STEP is one sharded build of the multi-phase library + the gather of its packed spot lists."""
    import torch
    import torch.distributed as dist
    from diffsims_b200 import SimulationGenerator
    from diffsims_b200.library import ShardedLibraryBuilder, active_quaternions
    w5 = WORKLOADS["c5"]
    H, W = WORKLOAD["shape"]
    B = args.batch or w5["batch"]
    per_phase = B * world // len(w5["phases"])
    phases5 = [(bench_phase(nm), active_quaternions(random_quats(per_phase, 10 + i))) for i, nm in enumerate(w5["phases"])]
    sb = ShardedLibraryBuilder(SimulationGenerator(w5["kv"]), phases5, rank=rank, world=world, reciprocal_radius=w5["rr"],
                               max_excitation_error=w5["s_max"], shape=(H, W), sigma=w5["sigma"], calibration=w5["rr"] / 128)
    sb.make_plan(render=True)
    for _ in range(max(1, args.warmup - 2)):
        sb.build(render=True)
        sb.gather()
        sb.result = None
    for b in sb.builders:
        b.launches = 0
    gather_ms = 0.0
    t0, t1 = _events()
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark_start()
        t0.record()
        for _ in range(args.steps):
            sb.result = None
            sb.build(render=True)
            a, b = _events()
            a.record()
            lib = sb.gather()
            b.record()
            torch.cuda.synchronize()
            gather_ms += a.elapsed_time(b)
        t1.record()
        barrier()
        clocks.mark_end()
    elapsed_ms = max_over_ranks(t0.elapsed_time(t1))
    n_total = per_phase * len(w5["phases"])
    launches = sum(b.launches for b in sb.builders)
    if rank == 0:
        line = dict(metric=METRIC, value=n_total * args.steps / (elapsed_ms * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=elapsed_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64+f32", data="synthetic",
                    config=dict(workload=w5["text"], workload_key="c5", orientations_total=n_total,
                                templates_per_rank=[sum(hi - lo for _, lo, hi in segs) for segs in sb.plan],
                                gather_ms_per_step=gather_ms / args.steps, gather_share=gather_ms / elapsed_ms,
                                gather_bytes_received_per_rank=int(sb.gather_bytes),
                                reflections_total=int(sum(int(p_.offsets[-1]) for p_ in lib))),
                    roofline=None, cpu_baseline=None, e2e=None, gpu_launches=launches, clocks=clocks.summary())
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="orientations per GPU per step (0: the workload's default)")
    ap.add_argument("--chunk", type=int, default=4096, help="e2e pipeline chunk (templates)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the per-config probes and the sharded multi-phase build")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels of a step one by one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from diffsims_b200 import SimulationGenerator, engine
    from diffsims_b200.library import ShardedLibraryBuilder, active_quaternions, gather_counts, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local_rank)   # pinned host buffers land on the GPU's own NUMA node
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))

    wname = args.workload
    w = WORKLOADS[wname]
    B, (H, W) = (args.batch or w["batch"]), WORKLOAD["shape"]
    if wname == "c5":
        return run_c5(args, rank, world, local_rank, dev, barrier, max_over_ranks, peak)

    # ================================================================================================
    # headline: one phase, rotation list sharded over the ranks (c5: see the sharded build below)
    # ================================================================================================
    ph0 = w["phases"][-1] if wname == "c5" else w["phases"][0]
    # weak scaling: the global rotation list has world * B entries, rank r owns a contiguous slice
    lo, hi = shard_bounds(world * B, rank, world)
    quats_all = random_quats(world * B, 0) if world * B <= (1 << 22) else None
    q_host = (quats_all[lo:hi] if quats_all is not None else random_quats(B, rank)).copy()
    q_host[:, 1:] *= -1  # active quaternions (the reference rotates g by ~rotation)
    q_dev = torch.as_tensor(q_host, device=dev)

    builder = make_builder(w, ph0, dev)
    gen = builder.gen
    builder.prepare()
    builder.calibrate_cap(q_dev)
    images = torch.empty((B, H, W), dtype=torch.float32, device=dev)

    k_events = {"k1": [], "k2": [], "k3": []}

    def step(timed):
        if timed:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
        builder.prepare()                       # K1 + table packing
        if timed:
            e[1].record()
        spots = builder.simulate(q_dev)         # K2
        if timed:
            e[2].record()
        builder.render(spots, images)           # K3 (+ its prepare pass on the tcgen05 path)
        if timed:
            e[3].record()
            for k, (a, b) in zip(("k1", "k2", "k3"), ((e[0], e[1]), (e[1], e[2]), (e[2], e[3]))):
                k_events[k].append((a, b))
        return spots

    for _ in range(args.warmup):
        step(False)
    barrier()
    # the timed loop replays ONE captured step (K1, pack, K2, K3 on fixed buffers); the kernels' own times are measured in
    # a separate eager loop because events cannot be recorded inside a graph
    graph = None
    if not args.no_graph:
        graph, spots = builder.capture(q_dev, images)
        for _ in range(args.warmup):
            graph.replay()
        barrier()
    launches_per_step = 3 + engine.render_launch_count(builder.cap, builder.shape, builder.sigma, True, builder.mean_spots)
    t0, t1 = _events()
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark_start()
        t0.record()
        for _ in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                spots = step(True)
        t1.record()
        barrier()
        clocks.mark_end()
    if graph is not None:   # kernel durations, same buffers, same stream, outside the graph
        for _ in range(min(args.steps, 10)):
            spots = step(True)
        torch.cuda.synchronize()
    elapsed_ms = max_over_ranks(t0.elapsed_time(t1))
    launches = launches_per_step * args.steps
    k_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in k_events.items()}
    # K1 / K2 are tens of microseconds: inside the eager step loop their event intervals also hold the host's launch gaps,
    # so they are timed on their own (median of back-to-back launches); K3's interval (>= 1 ms) is taken from the loop
    k_ms["k1"] = _time_ms(lambda: builder.prepare())
    k_ms["k2"] = _time_ms(lambda: builder.simulate(q_dev))
    builder.assert_no_overflow(spots)           # the unchecked timed passes stayed inside the calibrated capacity
    mean_spots = float(gather_counts(spots.count).float().mean().item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # context for the roofline: the pure-write ceiling of this GPU (the driver's peak is a read+write copy)
    fill_ms = []
    for _ in range(4):
        a, b = _events()
        a.record()
        images.zero_()
        b.record()
        torch.cuda.synchronize()
        fill_ms.append(a.elapsed_time(b))
    write_peak = B * H * W * 4 / (min(fill_ms[1:]) * 1e-3) / 1e9

    # ================================================================================================
    # end to end through the host-buffer calls
    # ================================================================================================
    e2e = None
    if not args.no_e2e:
        q_pin = torch.as_tensor(q_host).pin_memory()
        chunk = min(args.chunk, B)
        ring = [torch.empty((chunk, H, W), dtype=torch.float32).pin_memory() for _ in range(4)]
        keep = {}

        def consumer(lo_, hi_, view):     # the host side of the drop-in: here it only keeps the first templates for the check
            if lo_ == 0:
                keep["first"] = view[:64].clone()

        n_e2e = 3 if B >= 131072 else max(3, min(args.steps, 4))
        for _ in range(1 if B >= 131072 else 2):
            h2d, d2h = builder.run_host(q_pin, ring=ring, chunk=chunk, consumer=consumer)
        barrier()
        e0, e1 = _events()
        e0.record()
        for _ in range(n_e2e):
            h2d, d2h = builder.run_host(q_pin, ring=ring, chunk=chunk, consumer=consumer)   # (synchronises + capacity check)
        e1.record()
        barrier()
        te = max_over_ranks(e0.elapsed_time(e1))
        # correctness guard: the host images are the device images
        builder.render(spots, images)
        assert torch.equal(keep["first"], images[:64].cpu()), "e2e images differ from the device-resident images"
        e2e_value = world * B * n_e2e / (te * 1e-3)
        e2e = dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), steps=n_e2e,
                   api="TemplateLibraryBuilder.run_host (pinned host quaternions -> float32 images in a pinned host ring of "
                       f"{len(ring)} x {chunk} templates, consumer callback per chunk; capacity-checked)",
                   gbytes_per_s=e2e_value * H * W * 4 / 1e9, numa_node=numa)
        del ring
        # ---- the box's device->host roof, all ranks copying at once, no kernels
        roof_rank = d2h_roof(dev, world, barrier)
        roof = sum_over_ranks(roof_rank)
        e2e["d2h_roof"] = dict(gbytes_per_s_all_ranks=roof, gbytes_per_s_this_rank=roof_rank,
                               how="every rank: cudaMemcpyAsync of a 1 GiB device buffer into two pinned 1 GiB host slots for 1 s, "
                                   "concurrently, no kernels", e2e_fraction_of_roof=e2e["gbytes_per_s"] / roof)
        # ---- the optional 16-bit export of the same templates (rint(v * 65535); half the device->host bytes)
        ring16 = [torch.empty((chunk, H, W), dtype=torch.uint16).pin_memory() for _ in range(4)]
        builder.run_host(q_pin, ring=ring16, chunk=chunk)
        barrier()
        u0, u1 = _events()
        n_u = 2
        u0.record()
        for _ in range(n_u):
            h2d_u, d2h_u = builder.run_host(q_pin, ring=ring16, chunk=chunk)
        u1.record()
        barrier()
        tu = max_over_ranks(u0.elapsed_time(u1))
        e2e["e2e_uint16"] = dict(value=world * B * n_u / (tu * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d_u),
                                 d2h_bytes_per_step=int(d2h_u), steps=n_u, gbytes_per_s=world * B * n_u / (tu * 1e-3) * H * W * 2 / 1e9,
                                 api="TemplateLibraryBuilder.run_host into uint16 ring buffers: ds_render (float32) -> ds_quantize_u16 "
                                     "-> host; optional export, quantisation 7.6e-6 of the peak; float32 stays the headline")
        del ring16
        # ---- spot lists end to end: what calculate_diffraction2d returns (+ the polar arrays pyxem's matcher consumes)
        for _ in range(1):
            builder.run_host_spots(q_pin, polar=True)
        barrier()
        s0, s1 = _events()
        n_sp = 3
        s0.record()
        for _ in range(n_sp):
            packed, polar, h2d_s, d2h_s = builder.run_host_spots(q_pin, polar=True)
        s1.record()
        barrier()
        ts = max_over_ranks(s0.elapsed_time(s1))
        e2e["e2e_spots"] = dict(value=world * B * n_sp / (ts * 1e-3), unit=UNIT, h2d_bytes_per_step=int(h2d_s),
                                d2h_bytes_per_step=int(d2h_s), steps=n_sp, reflections=int(packed.offsets[-1]),
                                api="TemplateLibraryBuilder.run_host_spots(polar=True): pinned host quaternions -> K1, K2 "
                                    "(overflow-checked), ds_pack_csr, ds_polar_flatten -> packed CSR spot lists + padded "
                                    "(r, theta, intensity) arrays in pinned host memory")

    # ================================================================================================
    # BASELINE configs[4]: multi-phase library, sharded by cost, one gather at the end
    # ================================================================================================
    sharded = None
    extra = None
    del images
    torch.cuda.empty_cache()
    if not args.no_extra:
        w5 = WORKLOADS["c5"]
        per_phase = w5["batch"] * world // len(w5["phases"])
        phases5 = [(bench_phase(nm), active_quaternions(random_quats(per_phase, 10 + i))) for i, nm in enumerate(w5["phases"])]
        sb = ShardedLibraryBuilder(SimulationGenerator(w5["kv"]), phases5, rank=rank, world=world, reciprocal_radius=w5["rr"],
                                   max_excitation_error=w5["s_max"], shape=(H, W), sigma=w5["sigma"], calibration=w5["rr"] / 128)
        sb.make_plan(render=True)
        sb.build(render=True)       # warm-up (also calibrates the row capacities)
        lib = sb.gather()
        sb.result = lib = None      # (its images go back to the caching allocator: the timed build reuses the blocks)
        barrier()
        g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        g0.record()
        res = sb.build(render=True)
        g1.record()
        lib = sb.gather()
        g2.record()
        barrier()
        t_build, t_all = max_over_ranks(g0.elapsed_time(g1)), max_over_ranks(g0.elapsed_time(g2))
        n_total = per_phase * len(w5["phases"])
        sharded = dict(workload=w5["text"], orientations_total=n_total, phases=[nm for nm in w5["phases"]],
                       templates_per_rank=[sum(hi_ - lo_ for _, lo_, hi_ in segs) for segs in sb.plan],
                       value=n_total / (t_all * 1e-3), unit=UNIT, build_ms=t_build, gather_ms=t_all - t_build,
                       gather_share=(t_all - t_build) / t_all, gather_bytes_received_per_rank=int(sb.gather_bytes),
                       reflections_total=int(sum(int(p_.offsets[-1]) for p_ in lib)),
                       api="ShardedLibraryBuilder.build(render=True) + .gather(): K1 per phase on every rank, K2 + K3 + ds_pack_csr on "
                           "the rank's slice, all_gather of counts + padded all_gather of rows (images stay sharded)")
        del res, lib, sb
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the literal drop-in call sequence of the reference, objects and all (context for e2e) ---------
    if e2e is not None:
        from diffsims_b200.crystal import Rotation
        from diffsims_b200.generators.rotation_list_generators import get_beam_directions_grid
        grid = get_beam_directions_grid("cubic", 0.5)        # BASELINE configs[1] as written: 4 186 directions
        host_images = torch.empty((grid.shape[0], H, W), dtype=torch.float32).pin_memory()

        def drop_in():
            sim = gen.calculate_diffraction2d(si_phase(), Rotation.from_euler(grid, degrees=True),
                                              reciprocal_radius=WORKLOAD["rr"], with_direct_beam=True,
                                              max_excitation_error=WORKLOAD["s_max"])
            dev_images = sim.get_diffraction_patterns((H, W), sigma=WORKLOAD["sigma"],
                                                      calibration=WORKLOAD["calibration"])
            host_images.copy_(dev_images)
            return sim

        drop_in()
        torch.cuda.synchronize()
        t_api = time.perf_counter()
        for _ in range(3):
            drop_in()
        torch.cuda.synchronize()
        t_api = (time.perf_counter() - t_api) / 3
        e2e["drop_in_api"] = dict(
            value=grid.shape[0] / t_api, unit=UNIT, templates=int(grid.shape[0]), seconds=t_api,
            call="get_beam_directions_grid('cubic', 0.5) -> SimulationGenerator.calculate_diffraction2d -> "
                 "Simulation2D.get_diffraction_patterns -> host float32 (Python objects included, wall clock)")
        del host_images

    # ---- per-config probes (rank 0's GPU; the other ranks have left)
    if not args.no_extra and world == 1:
        extra = []
        for name in ("c2", "c2_dense", "c3", "c4", "c5"):
            extra.extend(config_probe(name, dev, peak))

    algo_bytes = B * H * W * 4
    k3_ms = k_ms["k3"]
    achieved = algo_bytes / (k3_ms * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "k3_traffic.json"
    if prof.exists():
        try:
            per_tmpl = json.loads(prof.read_text())["dram_bytes_per_template"]
            traffic = per_tmpl * B
        except Exception:
            traffic = None
    k3_kernels = engine.render_launch_count(builder.cap, builder.shape, builder.sigma, True, builder.mean_spots)
    # (the dispatch rule of ds_render: per-reflection tcgen05 product from 16 reflections per template, row-binned from 320)
    rows_kernel = k3_kernels == 2 and (builder.mean_spots or 0.0) >= 320.0
    roofline = dict(kernel="render_pipe_kernel (K3, ds_render)" if k3_kernels == 1 else
                           ("render_prepare_rows_kernel + render_rows_kernel (K3, ds_render, tcgen05 banded product)" if rows_kernel
                            else "render_prepare_kernel + render_umma_kernel (K3, ds_render, tcgen05)"),
                    bound="hbm", achieved=achieved, peak=peak, unit="GB/s",
                    frac=achieved / peak, traffic=traffic, algorithmic_bytes_per_launch=algo_bytes,
                    kernel_ms=k3_ms, peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    share_of_step=k3_ms * args.steps / elapsed_ms,
                    write_only_fill_gbs=write_peak, frac_of_write_only_fill=achieved / write_peak,
                    other_kernels=dict(k1_ms=k_ms["k1"], k2_ms=k_ms["k2"],
                                       k2_g_rows_per_s=B * builder.gtable.n / (k_ms["k2"] * 1e-3),
                                       note="K1 (structure factors + packing) and K2 (rotate / cull / refine) of the same step, "
                                            "CUDA events around eager launches"),
                    note="peak is the driver's read+write copy figure; a pure-write stream (torch fill of the same "
                         "buffer, write_only_fill_gbs) runs faster, so frac can exceed 1")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single_core()

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=elapsed_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64+f32", data="synthetic",
                config=dict(workload=w["text"], workload_key=wname, templates_per_gpu_per_step=B, n_g=int(builder.plan.hkl.shape[0]),
                            n_g_not_extinct=int(builder.gtable.n),
                            mean_spots_per_template=mean_spots, spot_capacity=int(builder.cap),
                            l2="each step writes %.1f GB of templates per GPU (>> 126 MB L2), so no input or "
                               "output survives in L2 between steps" % (algo_bytes / 1e9),
                            parallelism=f"rotation list sharded over {world} rank(s), no data-path collective",
                            launch=f"one CUDA graph replay per step ({launches_per_step} kernels)" if graph is not None
                                   else f"{launches_per_step} eager launches per step",
                            dtype_note="float64: structure factors, rotation, excitation error, shape factor, intensities, "
                                       "projection; float32: coarse cull and the rasterised templates"),
                roofline=roofline, cpu_baseline=cpu, e2e=e2e, e2e_sharded=sharded, extra=extra, gpu_launches=launches,
                clocks=clocks.summary())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
