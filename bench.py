#!/usr/bin/env python
"""Benchmark of the kinematical template-simulation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A STEP builds one template library of B orientations per GPU for the BASELINE config-2 workload
(Si diamond cubic, 200 kV, reciprocal_radius 1.0, max_excitation_error 0.01, lorentzian, lobato, direct
beam, 256 x 256 px float32 templates, sigma 10, calibration rr/128, normalised):
    K1 structure factors + table packing -> K2 simulate (all rotations) -> K3 rasterise (all templates).
`value` is templates/s over all GPUs with the rotation list already in HBM; `e2e` is the same library
built through the public host-buffer call (TemplateLibraryBuilder.run_host): pinned host quaternions in,
pinned host images out, every copy inside the timed region.  Rotation lists shard across ranks (weak
scaling: B per GPU) with no data-path collective.

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


def _cpu_init():
    from oracle import kinematical as K
    phase = si_phase()
    _CPU["K"] = K
    _CPU["phase"] = phase
    _CPU["gs"] = K.GSet(phase.structure, WORKLOAD["rr"], True)
    _CPU["wl"] = K.get_electron_wavelength(WORKLOAD["kv"])


def _cpu_templates(quats):
    """Spot list + rendered template for each quaternion with the float64 oracle; returns a checksum."""
    if not _CPU:
        _cpu_init()
    K = _CPU["K"]
    acc = 0.0
    for q in quats:
        G = K.quat_to_matrix(q)  # passive matrix: rotated = g @ G
        r = K.simulate_rotation(_CPU["phase"].structure, _CPU["gs"], G, _CPU["wl"], WORKLOAD["s_max"])
        img = K.diffraction_pattern(r["xyz"], r["intensity"], WORKLOAD["shape"], sigma=WORKLOAD["sigma"],
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
    per_worker = 24
    sample = cores * per_worker
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
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
                config=dict(workload=WORKLOAD["workload"], templates_per_step=sample,
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32768, help="orientations per GPU per step")
    ap.add_argument("--chunk", type=int, default=4096, help="e2e pipeline chunk (templates)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the four kernels of a step one by one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from diffsims_b200 import SimulationGenerator
    from diffsims_b200.library import TemplateLibraryBuilder, gather_counts, shard_bounds

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

    B, (H, W) = args.batch, WORKLOAD["shape"]
    # weak scaling: the global rotation list has world * B entries, rank r owns a contiguous slice
    lo, hi = shard_bounds(world * B, rank, world)
    quats_all = random_quats(world * B, 0) if world * B <= (1 << 22) else None
    q_host = (quats_all[lo:hi] if quats_all is not None else random_quats(B, rank)).copy()
    q_host[:, 1:] *= -1  # active quaternions (the reference rotates g by ~rotation)
    q_dev = torch.as_tensor(q_host, device=dev)

    gen = SimulationGenerator(WORKLOAD["kv"])
    builder = TemplateLibraryBuilder(gen, si_phase(), reciprocal_radius=WORKLOAD["rr"],
                                     max_excitation_error=WORKLOAD["s_max"], shape=(H, W), sigma=WORKLOAD["sigma"],
                                     calibration=WORKLOAD["calibration"])
    builder.prepare()
    builder.calibrate_cap(q_dev)
    images = torch.empty((B, H, W), dtype=torch.float32, device=dev)

    k3_events = []

    def step(timed):
        builder.prepare()                       # K1 + table packing
        spots = builder.simulate(q_dev)         # K2
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        builder.render(spots, images)           # K3
        if timed:
            b.record()
            k3_events.append((a, b))
        return spots

    for _ in range(args.warmup):
        step(False)
    barrier()
    # the timed loop replays ONE captured step (K1, pack, K2, K3 on fixed buffers); K3's own time is measured in a
    # separate eager loop because events cannot be recorded inside a graph
    graph = None
    if not args.no_graph:
        graph, spots = builder.capture(q_dev, images)
        for _ in range(args.warmup):
            graph.replay()
        barrier()
    builder.launches = 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark_start()
        t0.record()
        for _ in range(args.steps):
            if graph is not None:
                graph.replay()
                builder.launches += 4
            else:
                spots = step(True)
        t1.record()
        barrier()
        clocks.mark_end()
    if graph is not None:   # K3 launch duration, same buffers, same stream, outside the graph
        for _ in range(min(args.steps, 20)):
            step(True)
        torch.cuda.synchronize()
    elapsed_ms = t0.elapsed_time(t1)
    launches = 4 * args.steps if graph is not None else builder.launches
    k3_ms = float(np.mean([a.elapsed_time(b) for a, b in k3_events]))
    builder.assert_no_overflow(spots)           # the unchecked timed passes stayed inside the calibrated capacity
    all_counts = gather_counts(spots.count)     # the single collective of a sharded build (not timed)
    mean_spots = float(all_counts.float().mean().item())

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host-buffer call -----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        q_pin = torch.as_tensor(q_host).pin_memory()
        out_pin = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
        for _ in range(2):
            h2d, d2h = builder.run_host(q_pin, out_pin, chunk=args.chunk)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, min(args.steps, 4))
        e0.record()
        for _ in range(n_e2e):
            h2d, d2h = builder.run_host(q_pin, out_pin, chunk=args.chunk)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        builder.check_capacity()
        # correctness guard: the host images are the device images
        assert torch.equal(out_pin[:64], images[:64].cpu()), "e2e images differ from the device-resident images"
        e2e = dict(value=world * B * n_e2e / (float(te.item()) * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), steps=n_e2e,
                   api="TemplateLibraryBuilder.run_host (pinned host quaternions -> pinned host float32 images)",
                   numa_node=numa)

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

    # context for the roofline: the pure-write ceiling of this GPU (the driver's peak is a read+write copy)
    fill_ms = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        images.zero_()
        b.record()
        torch.cuda.synchronize()
        fill_ms.append(a.elapsed_time(b))
    write_peak = B * H * W * 4 / (min(fill_ms[1:]) * 1e-3) / 1e9

    peaks = {}
    for cand in (ROOT / "MEASURED_PEAKS.json",):
        if cand.exists():
            peaks = json.loads(cand.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    algo_bytes = B * H * W * 4
    achieved = algo_bytes / (k3_ms * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "k3_traffic.json"
    if prof.exists():
        try:
            per_tmpl = json.loads(prof.read_text())["dram_bytes_per_template"]
            traffic = per_tmpl * B
        except Exception:
            traffic = None
    roofline = dict(kernel="render_pipe_kernel (K3, ds_render)", bound="hbm", achieved=achieved, peak=peak, unit="GB/s",
                    frac=achieved / peak, traffic=traffic, algorithmic_bytes_per_launch=algo_bytes,
                    kernel_ms=k3_ms, peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    share_of_step=k3_ms * args.steps / elapsed_ms,
                    write_only_fill_gbs=write_peak,
                    note="peak is the driver's read+write copy figure; a pure-write stream (torch fill of the same "
                         "buffer, write_only_fill_gbs) runs faster, so frac can exceed 1")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single_core()

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=elapsed_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64+f32", data="synthetic",
                config=dict(workload=WORKLOAD["workload"], templates_per_gpu_per_step=B, n_g=int(builder.plan.hkl.shape[0]),
                            n_g_not_extinct=int(builder.gtable.n),
                            mean_spots_per_template=mean_spots, spot_capacity=int(builder.cap),
                            l2="each step writes %.1f GB of templates per GPU (>> 126 MB L2), so no input or "
                               "output survives in L2 between steps" % (algo_bytes / 1e9),
                            parallelism=f"rotation list sharded over {world} rank(s), no data-path collective",
                            launch="one CUDA graph replay per step (4 kernels)" if graph is not None else "4 eager launches per step",
                            dtype_note="float64: structure factors, rotation, excitation error, shape factor, intensities, "
                                       "projection; float32: coarse cull and the rasterised templates"),
                roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=launches, clocks=clocks.summary())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
