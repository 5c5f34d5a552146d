"""Batched template-library construction: the end-to-end call of the B200-native path.

The reference builds a library with a Python loop -- ``calculate_diffraction2d`` then one
``get_diffraction_pattern`` per template (simulation2d.py:357-442 renders only the current index).
``TemplateLibraryBuilder`` is the same computation for a whole rotation list of one phase:

    K1 (once per phase)  ->  K2 (all rotations of a chunk)  ->  K3 (all templates of a chunk)

* ``run_device``        inputs already on the device, images left on the device;
* ``run_host``          HOST rotations in, HOST float32 images out: chunks are pipelined over two CUDA streams so
                        the device->host copy of chunk i overlaps the kernels of chunk i+1.  The images land
                        either in a caller-provided pinned buffer of the whole library or in a small pinned RING
                        whose slots are handed to a consumer callback (a 300 k-template library is 79 GB);
* ``run_host_spots``    HOST rotations in, packed CSR spot lists on the HOST out -- what the reference's
                        ``calculate_diffraction2d`` returns (simulation_generator.py:244-259), ~40 B per reflection.

``ShardedLibraryBuilder`` is the multi-phase, multi-GPU form (library_generator.py:107-150 loops phases x
orientations): the (phase, rotation) units are split over the ranks in contiguous slices balanced by
sum N_rot * N_g, every rank builds its slice with no data-path collective, and ``gather()`` assembles the
packed spot lists on every rank with ONE all_gather of counts and ONE padded all_gather of rows.  Images
stay sharded behind the handle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi, engine

__all__ = ["TemplateLibraryBuilder", "ShardedLibraryBuilder", "PackedSpots", "shard_bounds", "split_work"]


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


@dataclass
class PackedSpots:
    """CSR spot lists of a rotation list: template t owns rows offsets[t] .. offsets[t+1]-1 (g-table order, i.e. the
    reference's reflection order).  Tensors live on one device (or on the host after ``.cpu()``)."""
    offsets: torch.Tensor     # [n + 1] int64
    g_index: torch.Tensor     # [total] int32, row of the phase's g table
    xyz: torch.Tensor         # [total, 3] float64
    intensity: torch.Tensor   # [total] float64

    @property
    def n(self):
        return self.offsets.shape[0] - 1

    def cpu(self):
        return PackedSpots(self.offsets.cpu(), self.g_index.cpu(), self.xyz.cpu(), self.intensity.cpu())

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.offsets, self.g_index, self.xyz, self.intensity))


def pack_csr(spots, total=None):
    """Device CSR packing of a padded ``engine.SpotTable`` (ds_pack_csr).  ``total`` (the sum of the counts) avoids
    the one synchronisation of this call when the caller already knows it."""
    dev = spots.count.device
    n = spots.n_rot
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(spots.count.clamp(max=spots.cap), 0, out=offsets[1:])
    if total is None:
        total = int(offsets[-1].item())
    g = torch.empty(total, dtype=torch.int32, device=dev)
    xyz = torch.empty((total, 3), dtype=torch.float64, device=dev)
    inten = torch.empty(total, dtype=torch.float64, device=dev)
    rc = _cabi.lib().ds_pack_csr(engine._stream(), n, spots.cap, _cabi.ptr(spots.count), _cabi.ptr(offsets),
                                 _cabi.ptr(spots.g_index), _cabi.ptr(spots.xyz), _cabi.ptr(spots.intensity),
                                 _cabi.ptr(g), _cabi.ptr(xyz), _cabi.ptr(inten))
    _cabi.check(rc, "ds_pack_csr")
    return PackedSpots(offsets, g, xyz, inten)


class CapacityError(RuntimeError):
    """A rotation produced more reflections than the packed rows hold (the extras were dropped)."""


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
        self.cap_is_checked = cap is not None  # an explicit capacity is the caller's promise; else calibrate_cap()
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

    def work_per_rotation(self, render=True):
        """Relative cost of one rotation of this phase, in picoseconds of B200 time (measured, profiles/): K2 culls the
        g table at ~0.6 ps per row and rotation plus ~1 ns per rotation (SURVEY.md section 8e balances by N_rot * N_g:
        that is the K2 term), and a rendered template costs its H * W * 4 bytes at the HBM write roof (40 ns at
        256 x 256) whatever the phase -- so rendered libraries split almost evenly by template count."""
        if self.gtable is None:
            self.prepare()
        cost = 1000 + 0.6 * int(self.gtable.n)
        if render:
            cost += self.shape[0] * self.shape[1] * 4 / 6.5536      # bytes / (6553.6 GB/s) in ps
        return max(1, int(cost))

    # -- K2 + K3 -------------------------------------------------------------------------------------
    def simulate(self, quats_dev, check_overflow=False):
        spots = engine.simulate(self.gtable, quats_dev, self.gen.wavelength, self.s_max, self.width, self.model,
                                self.minima, self.prec, self.gen.minimum_intensity, cap=self.cap,
                                check_overflow=check_overflow)
        self.launches += 1
        self.cap = spots.cap
        return spots

    def render(self, spots, out):
        self.launches += engine.render_launch_count(spots.cap, self.shape, self.sigma, self.fast, self.mean_spots)
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
        self.cap_is_checked = True
        return self.cap

    def assert_no_overflow(self, spots):
        """Host check (synchronises) that an unchecked pass did not outgrow the capacity."""
        if spots.max_count is not None and int(spots.max_count.item()) > spots.cap:
            raise CapacityError(f"{int(spots.max_count.item())} reflections in one rotation exceed the capacity "
                                f"{spots.cap}: call calibrate_cap() on this rotation list")

    def _require_checked_cap(self, what):
        if not self.cap_is_checked:
            raise RuntimeError(f"{what} launches K2 without a capacity check (it must not synchronise): call "
                               "calibrate_cap() on a representative rotation list first, or pass cap= explicitly. "
                               "An under-sized capacity would silently truncate dense orientations.")

    def run_device(self, quats_dev, out_images):
        """Device-resident pass: K1 + K2 + K3 on the current stream, no synchronisation; returns the SpotTable.
        Needs a calibrated (or explicitly given) row capacity; ``assert_no_overflow(spots)`` verifies it afterwards."""
        self.prepare()
        self._require_checked_cap("run_device")
        spots = self.simulate(quats_dev)
        self.render(spots, out_images)
        return spots

    def capture(self, quats_dev, out_images):
        """Capture one device-resident library build (K1 + pack + K2 + K3 on fixed buffers) into a CUDA graph.
        Returns (graph, spots): ``graph.replay()`` re-runs the build with whatever ``quats_dev`` then holds,
        without per-launch host work -- the kernels of a sparse library take ~1.3 ms per 32 k templates, so the few
        microseconds between dependent launches are worth removing."""
        self._require_checked_cap("capture")
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

    # -- host-buffer paths ---------------------------------------------------------------------------
    def run_host(self, quats_host, out_host=None, chunk=8192, counts_host=None, ring=None, consumer=None):
        """HOST buffers in and out.  ``quats_host``: pinned float64 tensor [n, 4] (active quaternions).

        Images go either to ``out_host`` (pinned float32 [n, H, W], the whole library) or through ``ring``: a list
        of >= 2 pinned float32 tensors [chunk, H, W]; chunk i lands in ``ring[i % len(ring)]`` and, once its copy has
        completed, ``consumer(lo, hi, host_view)`` is called (the slot is reused ``len(ring)`` chunks later, so the
        consumer must be done with the view when it returns).  The call synchronises at its end and RAISES
        ``CapacityError`` if any rotation produced more reflections than the rows hold (nothing is returned from a
        truncated build).  Returns (h2d_bytes, d2h_bytes).

        Destinations of dtype ``torch.uint16`` receive the optional 16-bit export (``engine.quantize_u16``: normalised
        templates as rint(v * 65535), half the device->host bytes); float32 is the default and the parity-tested form."""
        if (out_host is None) == (ring is None):
            raise ValueError("run_host needs exactly one of out_host= and ring=")
        as_u16 = (ring[0] if ring is not None else out_host).dtype == torch.uint16
        if as_u16 and not self.normalize:
            raise ValueError("the uint16 export needs normalised templates (values in [0, 1])")
        dev = engine.device()
        n = quats_host.shape[0]
        H, W = self.shape
        self.prepare()
        if ring is not None:
            chunk = min(chunk, ring[0].shape[0])
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        bufs = [torch.empty((min(chunk, n), H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        bufs16 = [torch.empty((min(chunk, n), H, W), dtype=torch.uint16, device=dev) for _ in range(2)] if as_u16 else None
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        h2d = d2h = 0
        worst = torch.zeros(2, dtype=torch.int32, device=dev)   # per stream: largest pre-cut reflection count
        pending = []                                            # ring mode: (event, lo, hi, slot) not yet consumed
        for i, lo in enumerate(range(0, n, chunk)):
            hi = min(lo + chunk, n)
            st, buf = streams[i & 1], bufs[i & 1][: hi - lo]
            if ring is not None:
                slot = i % len(ring)
                while pending and (pending[0][3] == slot or len(pending) >= len(ring)):
                    ev, plo, phi, pslot = pending.pop(0)      # the slot's previous occupant must have been consumed
                    ev.synchronize()
                    if consumer is not None:
                        consumer(plo, phi, ring[pslot][: phi - plo])
                dst = ring[slot][: hi - lo]
            else:
                dst = out_host[lo:hi]
            st.wait_event(ready)  # the g table is produced on the caller's stream
            with torch.cuda.stream(st):
                q = quats_host[lo:hi].to(dev, non_blocking=True)
                spots = self.simulate(q)
                self.render(spots, buf)
                w = worst[(i & 1):(i & 1) + 1]
                torch.maximum(w, spots.max_count, out=w)
                if as_u16:
                    dst.copy_(engine.quantize_u16(buf, bufs16[i & 1][: hi - lo]), non_blocking=True)
                    self.launches += 1
                else:
                    dst.copy_(buf, non_blocking=True)
                if counts_host is not None:
                    counts_host[lo:hi].copy_(spots.count, non_blocking=True)
                    d2h += (hi - lo) * 4
                if ring is not None:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    pending.append((ev, lo, hi, slot))
            h2d += (hi - lo) * 32
            d2h += (hi - lo) * H * W * (2 if as_u16 else 4)
        for st in streams:
            main.wait_stream(st)
        for ev, plo, phi, pslot in pending:
            ev.synchronize()
            if consumer is not None:
                consumer(plo, phi, ring[pslot][: phi - plo])
        self.last_max_count = worst
        self.check_capacity()     # synchronises; a truncated build raises instead of returning
        return h2d, d2h

    def run_host_spots(self, quats_host, chunk=131072, polar=False):
        """HOST rotations in, packed spot lists on the HOST out: ``PackedSpots`` in pinned memory (and, with
        ``polar=True``, the padded (r, theta, intensity) arrays of polar_flatten_simulations, one triple per chunk).
        K1 -> K2 -> CSR pack -> device->host copy of ~40 bytes per reflection; no image is rendered.  One host
        synchronisation per chunk (the row total and the capacity check travel together); a chunk that outgrew the
        row capacity is redone with a larger one.  The pinned result buffers belong to the builder and are reused by
        the next call.  Returns (packed, h2d_bytes, d2h_bytes) or (packed, polar_arrays, h2d_bytes, d2h_bytes)."""
        dev = engine.device()
        n = quats_host.shape[0]
        self.prepare()
        parts, polar_parts = [], []
        h2d = d2h = 0
        pool = self.__dict__.setdefault("_pinned_pool", {})

        def pinned(key, shape, dtype):
            need = int(np.prod(shape))
            t = pool.get(key)
            if t is None or t.numel() < need or t.dtype != dtype:
                t = pool[key] = torch.empty(max(need, 1), dtype=dtype, pin_memory=True)
            return t[:need].view(shape)

        word = pinned("word", (3,), torch.int64)
        for ci, lo in enumerate(range(0, n, chunk)):
            hi = min(lo + chunk, n)
            q = quats_host[lo:hi].to(dev, non_blocking=True)
            while True:
                spots = self.simulate(q)                      # unchecked launch; verified with the row total below
                offsets = torch.zeros(hi - lo + 1, dtype=torch.int64, device=dev)
                torch.cumsum(spots.count.clamp(max=spots.cap), 0, out=offsets[1:])
                stat = torch.stack([offsets[-1], spots.max_count[0].to(torch.int64), spots.count.max().to(torch.int64)])
                word.copy_(stat, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                total, need, longest = (int(v) for v in word)
                if need <= spots.cap:
                    break
                self.cap = (need + 31) // 32 * 32             # a denser orientation than the capacity allowed: redo
            g = torch.empty(total, dtype=torch.int32, device=dev)
            xyz = torch.empty((total, 3), dtype=torch.float64, device=dev)
            inten = torch.empty(total, dtype=torch.float64, device=dev)
            _cabi.check(_cabi.lib().ds_pack_csr(engine._stream(), hi - lo, spots.cap, _cabi.ptr(spots.count), _cabi.ptr(offsets),
                                                _cabi.ptr(spots.g_index), _cabi.ptr(spots.xyz), _cabi.ptr(spots.intensity),
                                                _cabi.ptr(g), _cabi.ptr(xyz), _cabi.ptr(inten)), "ds_pack_csr")
            self.launches += 1
            host = PackedSpots(pinned(("off", ci), offsets.shape, torch.int64), pinned(("g", ci), g.shape, torch.int32),
                               pinned(("xyz", ci), xyz.shape, torch.float64), pinned(("I", ci), inten.shape, torch.float64))
            for d, s_ in zip((host.offsets, host.g_index, host.xyz, host.intensity), (offsets, g, xyz, inten)):
                d.copy_(s_, non_blocking=True)
            parts.append(host)
            d2h += host.nbytes() + 24
            if polar:
                r, t, i = engine.polar_flatten(spots.count, spots.xyz, spots.intensity, max(longest, 1))
                self.launches += 1
                hp = [pinned((nm, ci), x.shape, x.dtype) for nm, x in zip(("pr", "pt", "pi"), (r, t, i))]
                for d, s_ in zip(hp, (r, t, i)):
                    d.copy_(s_, non_blocking=True)
                polar_parts.append(hp)
                d2h += sum(x.numel() * 8 for x in hp)
            h2d += (hi - lo) * 32
        torch.cuda.current_stream().synchronize()
        packed = concat_packed(parts)
        if polar:
            return packed, polar_parts, h2d, d2h
        return packed, h2d, d2h

    def check_capacity(self):
        """After ``run_host`` (synchronises): raise if a chunk produced more reflections than the rows hold."""
        worst = int(self.last_max_count.max().item())
        if worst > self.cap:
            raise CapacityError(f"{worst} reflections in one rotation exceed the capacity {self.cap}: "
                                f"call calibrate_cap() on this rotation list")


def concat_packed(parts):
    """Concatenate CSR pieces that follow each other in template order."""
    if len(parts) == 1:
        return parts[0]
    offs, base = [parts[0].offsets], int(parts[0].offsets[-1])
    for p in parts[1:]:
        offs.append(p.offsets[1:] + base)
        base += int(p.offsets[-1])
    return PackedSpots(torch.cat(offs), torch.cat([p.g_index for p in parts]), torch.cat([p.xyz for p in parts]),
                       torch.cat([p.intensity for p in parts]))


# ------------------------------------------------------------------------------------------------------------------
# multi-phase, multi-GPU
# ------------------------------------------------------------------------------------------------------------------
def split_work(counts, costs, world):
    """Split the concatenated (phase, rotation) units -- ``counts[p]`` rotations of cost ``costs[p]`` each, phase
    after phase as the reference loops them (library_generator.py:107, :117) -- into ``world`` contiguous slices of
    (nearly) equal total cost.  Returns, per rank, a list of (phase, lo, hi) segments.  Deterministic and identical
    on every rank (pure integer arithmetic on the inputs)."""
    counts = [int(c) for c in counts]
    costs = [max(1, int(c)) for c in costs]
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)                    # first unit of each phase
    start = np.concatenate([[0], np.cumsum([c * w for c, w in zip(counts, costs)])]).astype(np.int64)
    total, n_units = int(start[-1]), int(first[-1])

    def unit_at(target):
        """Index of the first unit whose cumulative cost (before it) is >= target."""
        p = int(np.searchsorted(start, target, side="right")) - 1
        if p >= len(counts):
            return n_units
        k = -(-(target - int(start[p])) // costs[p])          # ceil
        return int(first[p]) + min(k, counts[p])

    bounds = [0] + [unit_at((total * r) // world) for r in range(1, world)] + [n_units]
    bounds = [min(max(b, 0), n_units) for b in bounds]
    for r in range(1, len(bounds)):
        bounds[r] = max(bounds[r], bounds[r - 1])
    out = []
    for r in range(world):
        u0, u1 = bounds[r], bounds[r + 1]
        segs = []
        for p in range(len(counts)):
            lo, hi = max(u0, int(first[p])) - int(first[p]), min(u1, int(first[p + 1])) - int(first[p])
            if hi > lo:
                segs.append((p, int(lo), int(hi)))
        out.append(segs)
    return out


@dataclass
class ShardResult:
    """What one rank holds after ``ShardedLibraryBuilder.build``: per segment the packed spot lists (device) and, if
    rendered, the images of its templates (device; they stay sharded)."""
    segments: list            # [(phase, lo, hi)]
    packed: list              # [PackedSpots] per segment
    images: list              # [tensor [hi - lo, H, W] float32 or None] per segment


class ShardedLibraryBuilder:
    """Multi-phase template library over the ranks of ``torch.distributed`` (one process per GPU).

    ``phases``: list of (phase, quats) with ``quats`` the ACTIVE quaternions [n_p, 4] (numpy or tensor) of the
    phase's orientation list; every rank passes the same lists (they are tiny compared with the result: 32 B per
    orientation) and takes its slice.  ``builder_kwargs`` go to every per-phase ``TemplateLibraryBuilder``."""

    def __init__(self, generator, phases, rank=None, world=None, **builder_kwargs):
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank, self.world = int(rank), int(world)
        self.phases = phases
        self.builders = [TemplateLibraryBuilder(generator, ph, **builder_kwargs) for ph, _ in phases]
        self.counts = [int(np.asarray(q.shape)[0]) for _, q in phases]
        self.result = None
        self.plan = None

    def make_plan(self, render=True):
        """K1 for every phase (each rank computes the tiny F(g) tables locally: no broadcast needed) and the split."""
        costs = [b.work_per_rotation(render) for b in self.builders]
        self.plan = split_work(self.counts, costs, self.world)
        return self.plan

    def build(self, render=True, calibrate=4096):
        """Build this rank's slice.  Row capacities are calibrated per phase on (a sample of) the rank's own slice and
        every pass is overflow-checked afterwards (CapacityError).  Returns the ``ShardResult``."""
        dev = engine.device()
        if self.plan is None:
            self.make_plan(render)
        segs = self.plan[self.rank]
        tables, images = [], []
        for p, lo, hi in segs:
            b = self.builders[p]
            q = torch.as_tensor(np.ascontiguousarray(np.asarray(self.phases[p][1])[lo:hi], dtype=np.float64), device=dev)
            if not b.cap_is_checked:
                b.calibrate_cap(q[: min(calibrate, hi - lo)])
            spots = b.simulate(q)
            img = None
            if render:
                img = torch.empty((hi - lo, *b.shape), dtype=torch.float32, device=dev)
                b.render(spots, img)
            tables.append((b, spots))
            images.append(img)
        packed = []
        for b, spots in tables:      # (the first .item() synchronises once, after everything has been enqueued)
            need = int(spots.max_count.item())
            if need > spots.cap:      # a dense orientation outside the calibration sample
                raise CapacityError(f"{need} reflections in one rotation exceed the capacity {spots.cap} of phase "
                                    f"{getattr(b.phase, 'name', '?')}: raise `calibrate` or pass cap=")
            packed.append(pack_csr(spots))
            b.launches += 1
        self.result = ShardResult(segs, packed, images)
        return self.result

    # -- the one exchange of a sharded build ---------------------------------------------------------------
    def gather(self, group=None):
        """Assemble the packed spot lists of the WHOLE library on every rank (``gather_shards``).  Returns a list with
        one ``PackedSpots`` per phase, in the phase's orientation order -- bit-identical to a single-rank build."""
        out, self.gather_bytes = gather_shards(self.plan, self.result.packed, self.rank, self.world, len(self.phases),
                                               group=group)
        return out


def gather_shards(plan, packed, rank, world, n_phases, group=None):
    """The one exchange of a sharded build: ``plan`` (split_work) says which (phase, lo, hi) segments every rank
    built, ``packed`` holds this rank's ``PackedSpots`` per segment.  One all_gather of the per-template counts
    (padded to the longest shard) and one padded all_gather of the rows (x, y, z, intensity, g index as five
    float64 -- the index is exact) put the whole library on every rank.  Works on any backend / device the tensors
    live on (NCCL on GPUs, gloo on the host).  Returns ([PackedSpots per phase], bytes received per rank)."""
    import torch.distributed as dist
    dev = packed[0].offsets.device if packed else torch.device("cpu")
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
    counts_local = torch.cat([p.offsets[1:] - p.offsets[:-1] for p in packed]) if packed else z(0, torch.int64)
    g_local = torch.cat([p.g_index for p in packed]) if packed else z(0, torch.int32)
    x_local = torch.cat([p.xyz for p in packed]) if packed else z((0, 3), torch.float64)
    i_local = torch.cat([p.intensity for p in packed]) if packed else z(0, torch.float64)
    nbytes = 0
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        all_counts, rows = [counts_local], [(g_local, x_local, i_local)]
    else:
        n_units = [sum(hi - lo for _, lo, hi in segs) for segs in plan]   # every rank knows every shard's size
        m = max(max(n_units), 1)
        pad = z(m, torch.int64)
        pad[: counts_local.numel()] = counts_local
        got = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(got, pad, group=group)
        all_counts = [got[r][: n_units[r]] for r in range(world)]
        totals = [int(c.sum().item()) for c in all_counts]
        mr = max(max(totals), 1)
        payload = z((mr, 5), torch.float64)
        k = g_local.numel()
        payload[:k, :3] = x_local
        payload[:k, 3] = i_local
        payload[:k, 4] = g_local.to(torch.float64)
        got_rows = [torch.empty_like(payload) for _ in range(world)]
        dist.all_gather(got_rows, payload, group=group)
        nbytes = world * (m * 8 + mr * 40)
        rows = []
        for r in range(world):
            blk = got_rows[r][: totals[r]]
            rows.append((blk[:, 4].to(torch.int32), blk[:, :3].contiguous(), blk[:, 3].contiguous()))
    # stitch the rank shards back into per-phase lists
    per_phase = [[] for _ in range(n_phases)]
    ranks = range(world) if len(all_counts) == world else [rank]
    for r, c, (g, x, i) in zip(ranks, all_counts, rows):
        cu, ru = 0, 0      # cursors over this rank's units / rows
        for p, lo, hi in plan[r]:
            cnt = c[cu: cu + (hi - lo)]
            nrow = int(cnt.sum().item())
            off = z(hi - lo + 1, torch.int64)
            torch.cumsum(cnt, 0, out=off[1:])
            per_phase[p].append(PackedSpots(off, g[ru: ru + nrow], x[ru: ru + nrow], i[ru: ru + nrow]))
            cu += hi - lo
            ru += nrow
    empty = lambda: PackedSpots(z(1, torch.int64), z(0, torch.int32), z((0, 3), torch.float64), z(0, torch.float64))
    return [concat_packed(parts) if parts else empty() for parts in per_phase], nbytes


def gather_counts(local_counts):
    """Every rank learns all per-template spot counts (kept for callers that only need the counts)."""
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
