"""Batched template-library construction: the end-to-end call of the B200-native path.

The reference builds a library with a Python loop -- ``calculate_diffraction2d`` then one
``get_diffraction_pattern`` per template (simulation2d.py:357-442 renders only the current index).
``TemplateLibraryBuilder`` is the same computation for a whole rotation list:

    K1 (once per phase)  ->  K2 (all rotations of a chunk)  ->  K3 (all templates of a chunk)

* ``run_device``  inputs already on the device, images left on the device;
* ``run_host``    HOST rotations in, HOST images out: chunks are pipelined over two CUDA streams so the
                  device->host copy of chunk i overlaps the kernels of chunk i+1;
* rotation lists shard across ranks with no data-path collective (every (phase, rotation) unit is
  independent, simulation_generator.py:198/:211); ``gather_counts`` is the single collective at the end.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .crystal import Rotation

__all__ = ["TemplateLibraryBuilder", "shard_bounds"]


def shard_bounds(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of ``n`` rotations owned by ``rank``."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def active_quaternions(rotation):
    """Quaternions the kernel applies: the reference rotates g by ``~rotation``
    (_diffracting_vector.py:160).  Accepts an orix-like Rotation or an [n, 4] array of (a, b, c, d)."""
    q = np.asarray(getattr(rotation, "data", rotation), dtype=np.float64).reshape(-1, 4).copy()
    q[:, 1:] *= -1.0
    return q


class TemplateLibraryBuilder:
    def __init__(self, generator, phase, reciprocal_radius=1.0, with_direct_beam=True,
                 max_excitation_error=1e-2, shape_factor_width=None, debye_waller_factors=None,
                 shape=(256, 256), sigma=10, calibration=0.01, direct_beam_position=None, in_plane_angle=0,
                 mirrored=False, fast=True, normalize=True, clip_threshold=1, cap=None):
        self.gen = generator
        self.phase = phase
        self.rr = reciprocal_radius
        self.with_direct_beam = with_direct_beam
        self.s_max = max_excitation_error
        self.width = max_excitation_error if shape_factor_width is None else shape_factor_width
        self.dw = debye_waller_factors or {}
        self.shape = (int(shape[0]), int(shape[1]))
        self.sigma, self.calibration = sigma, calibration
        if direct_beam_position is None:
            H, W = self.shape
            direct_beam_position = (W // 2, H // 2) if fast else ((W - 1) / 2, (H - 1) / 2)
        self.center = direct_beam_position
        self.angle, self.mirrored, self.fast = in_plane_angle, mirrored, fast
        self.normalize, self.clip = normalize, clip_threshold
        self.model, self.minima = generator._native_model()
        if self.model is None:
            raise NotImplementedError("TemplateLibraryBuilder needs a native shape factor model")
        self.prec = float(np.deg2rad(generator.precession_angle))
        self.cap = cap
        self.gtable = None
        self.plan = None   # host enumeration + uploads of the phase, made once
        self.launches = 0  # kernels of libdiffsims_b200.so launched by this builder
        self.mean_spots = None  # mean reflections per template seen by calibrate_cap (schedule hint for K3)

    # -- K1 ----------------------------------------------------------------------------------------
    def prepare(self):
        """K1: structure factors of the phase's g set + the float4 table K2 stages (two launches, async).
        The integer hkl enumeration and the atom table are a host-side plan built on first use."""
        if self.plan is None:
            self.plan = self.gen._g_plan(self.phase, self.rr, self.with_direct_beam, self.dw)
        self.gtable = self.plan.run(self.gen._extinct_rel_cut(self.with_direct_beam))
        self.launches += 2  # structure factors + table packing
        if self.cap is None:
            self.cap = engine.estimate_cap(self.gtable.n, self.gtable.g_max, self.s_max, self.prec)
        return self.gtable

    # -- K2 + K3 -------------------------------------------------------------------------------------
    def simulate(self, quats_dev, check_overflow=False):
        spots = engine.simulate(self.gtable, quats_dev, self.gen.wavelength, self.s_max, self.width, self.model,
                                self.minima, self.prec, self.gen.minimum_intensity, cap=self.cap,
                                check_overflow=check_overflow)
        self.launches += 1
        self.cap = spots.cap
        return spots

    def render(self, spots, out):
        self.launches += 1
        return engine.render(spots.count, spots.xyz, spots.intensity, self.shape, self.sigma, self.calibration,
                             self.center, self.angle, self.mirrored, self.fast, self.normalize, self.clip, out=out,
                             mean_spots=self.mean_spots)

    def calibrate_cap(self, quats_dev):
        """One untimed overflow-checked pass that fixes ``cap`` for the rotation list."""
        spots = self.simulate(quats_dev, check_overflow=True)
        # the rows must hold every reflection that passes the excitation-error cut, i.e. the count BEFORE the
        # minimum-intensity cut (K2 compacts in place); the later unchecked passes rely on this capacity
        need = int(spots.max_count.item()) if spots.n_rot and spots.max_count is not None else 0
        self.cap = max(32, (max(need, 1) + 31) // 32 * 32)
        self.mean_spots = float(spots.count.float().mean().item()) if spots.n_rot else None
        return self.cap

    def assert_no_overflow(self, spots):
        """Host check (synchronises) that an unchecked pass did not outgrow the calibrated capacity."""
        if spots.max_count is not None and int(spots.max_count.item()) > spots.cap:
            raise RuntimeError(f"{int(spots.max_count.item())} reflections in one rotation exceed the capacity "
                               f"{spots.cap}: call calibrate_cap() on this rotation list")

    def run_device(self, quats_dev, out_images):
        """Device-resident pass: K1 + K2 + K3 on the current stream; returns the SpotTable."""
        self.prepare()
        spots = self.simulate(quats_dev)
        self.render(spots, out_images)
        return spots

    def capture(self, quats_dev, out_images):
        """Capture one device-resident library build (K1 + pack + K2 + K3 on fixed buffers) into a CUDA graph.
        Returns (graph, spots): ``graph.replay()`` re-runs the build with whatever ``quats_dev`` then holds,
        without per-launch host work -- the four kernels of a sparse library take ~1.3 ms, so the few
        microseconds between dependent launches are worth removing."""
        assert self.cap is not None, "call calibrate_cap() first: the captured buffers have a fixed capacity"
        side = torch.cuda.Stream(device=quats_dev.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up on the capture stream (lazy module loading, attributes)
            self.run_device(quats_dev, out_images)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            spots = self.run_device(quats_dev, out_images)
        return graph, spots

    def run_host(self, quats_host, out_host, chunk=8192, counts_host=None):
        """HOST buffers in and out.  ``quats_host``: pinned float64 tensor [n, 4] (active quaternions);
        ``out_host``: pinned float32 tensor [n, H, W].  Returns (h2d_bytes, d2h_bytes)."""
        dev = engine.device()
        n = quats_host.shape[0]
        H, W = self.shape
        self.prepare()
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        bufs = [torch.empty((min(chunk, n), H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        h2d = d2h = 0
        worst = torch.zeros(2, dtype=torch.int32, device=dev)   # per stream: largest pre-cut reflection count
        for i, lo in enumerate(range(0, n, chunk)):
            hi = min(lo + chunk, n)
            st, buf = streams[i & 1], bufs[i & 1][: hi - lo]
            st.wait_event(ready)  # the g table is produced on the caller's stream
            with torch.cuda.stream(st):
                q = quats_host[lo:hi].to(dev, non_blocking=True)
                spots = self.simulate(q)
                self.render(spots, buf)
                w = worst[(i & 1):(i & 1) + 1]
                torch.maximum(w, spots.max_count, out=w)
                out_host[lo:hi].copy_(buf, non_blocking=True)
                if counts_host is not None:
                    counts_host[lo:hi].copy_(spots.count, non_blocking=True)
                    d2h += (hi - lo) * 4
            h2d += (hi - lo) * 32
            d2h += (hi - lo) * H * W * 4
        for st in streams:
            main.wait_stream(st)
        self.last_max_count = worst   # device word; compare with self.cap after synchronising (check_capacity)
        return h2d, d2h

    def check_capacity(self):
        """After ``run_host`` (synchronises): raise if a chunk produced more reflections than the rows hold."""
        worst = int(self.last_max_count.max().item())
        if worst > self.cap:
            raise RuntimeError(f"{worst} reflections in one rotation exceed the capacity {self.cap}: "
                               f"call calibrate_cap() on this rotation list")


def gather_counts(local_counts):
    """The one collective of a sharded library build: every rank learns all per-template spot counts."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_counts
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_counts.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_counts.numel()], dtype=torch.int64, device=local_counts.device))
    m = int(max(s.item() for s in sizes))
    pad = torch.zeros(m, dtype=local_counts.dtype, device=local_counts.device)
    pad[: local_counts.numel()] = local_counts
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(out, sizes)])
