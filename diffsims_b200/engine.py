"""Device-level plumbing between the Python API mirror and the C-ABI kernels.

torch is used for device memory, streams and host<->device copies only; every
arithmetic stage of the hot path runs in libdiffsims_b200.so (K1 structure factors,
K2 simulate, K3 render).  There is no CPU fallback.
"""
from __future__ import annotations

import json
import math
import warnings
from dataclasses import dataclass
from pathlib import Path

import os

import numpy as np
import torch

from . import _cabi
from .crystal import get_element

SHAPE_MODEL_IDS = {"binary": 0, "linear": 1, "sinc": 2, "sin2c": 3, "atanc": 4, "lorentzian": 5,
                   "lorentzian_precession": 6, "return_s": 7}
SCATTERING_IDS = {None: 0, "lobato": 1, "xtables": 2}

_TABLES = None


def scattering_tables():
    global _TABLES
    if _TABLES is None:
        p = Path(__file__).resolve().parent / "data" / "scattering_params.json"
        _TABLES = json.loads(p.read_text())
    return _TABLES


def get_scattering_params_dict(scattering_params):
    """diffsims/utils/sim_utils.py:139-162."""
    if scattering_params in ("lobato", "xtables"):
        return scattering_tables()[scattering_params]
    raise NotImplementedError(
        "The scattering parameters `{}` are not implemented. "
        "See documentation for available implementations.".format(scattering_params))


def device(dev=None):
    if not torch.cuda.is_available():
        raise RuntimeError("diffsims_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    _cabi.lib()
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(dev)


def _stream():
    return _cabi.c_void_p(torch.cuda.current_stream().cuda_stream)


# ----------------------------------------------------------------------------------------------
# K1
# ----------------------------------------------------------------------------------------------
def atom_arrays(structure, debye_waller_factors, scattering_params):
    """Host-side flattening of a structure, grouped by element.

    Mirrors get_vectorized_list_for_atomic_scattering_factors (diffsims/utils/sim_utils.py:165-224)
    -- including the warning + zero coefficients for unknown elements -- and the change of
    reference frame of the fractional coordinates at :290-291.
    """
    if debye_waller_factors is None:
        debye_waller_factors = {}
    tbl = get_scattering_params_dict(scattering_params) if scattering_params is not None else {}
    lat = structure.lattice
    mat = np.linalg.inv(np.dot(np.asarray(lat.stdbase, float), np.asarray(lat.recbase, float)))
    elements, groups = [], {}
    for site in structure:
        el = get_element(site.element)
        if el not in groups:
            groups[el] = []
            elements.append(el)
            if el not in tbl:
                warnings.warn(f"Element {el} from atom type symbol {site.element} not "
                              "found in scattering parameter library.")
        groups[el].append(site)
    frac, occ, start, coeffs, dw = [], [], [0], [], []
    for el in elements:
        for site in groups[el]:
            frac.append(np.dot(np.asarray(site.xyz, float), mat))
            occ.append(float(site.occupancy))
        start.append(len(frac))
        coeffs.append(np.asarray(tbl.get(el, np.zeros(10)), float).reshape(10))
        dw.append(float(debye_waller_factors.get(el, 0)))
    return (np.asarray(frac, float).reshape(-1, 3), np.asarray(occ, float), np.asarray(start, np.int32),
            np.asarray(coeffs, float).reshape(-1, 10), np.asarray(dw, float))


class AtomTable:
    """Device copy of a structure's atom arrays (grouped by element) for K1."""

    def __init__(self, structure, debye_waller_factors, scattering_params, dev):
        if scattering_params not in SCATTERING_IDS:
            raise NotImplementedError(
                "The scattering parameters `{}` are not implemented. "
                "See documentation for available implementations.".format(scattering_params))
        frac, occ, start, coeffs, dw = atom_arrays(structure, debye_waller_factors, scattering_params)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        self.n_atoms, self.n_elem = frac.shape[0], coeffs.shape[0]
        self.frac, self.occ, self.start, self.coeffs, self.dw = t(frac), t(occ), t(start), t(coeffs), t(dw)
        self.model = SCATTERING_IDS[scattering_params]


def launch_structure_factors(atoms: AtomTable, hkl_d, gnorm_d, prefactor_d=None, F=None, I=None, hkl_int_max=0,
                             scratch=None):
    """Enqueue K1 on the current stream (no host work, no synchronisation).  ``hkl_int_max`` > 0 (all indices are
    integers of at most that magnitude) with a ``scratch`` tensor lets large cells take the factorised kernels."""
    rc = _cabi.lib().ds_structure_factors(
        _stream(), hkl_d.shape[0], _cabi.ptr(hkl_d), _cabi.ptr(gnorm_d), atoms.n_atoms, _cabi.ptr(atoms.frac),
        _cabi.ptr(atoms.occ), atoms.n_elem, _cabi.ptr(atoms.start), _cabi.ptr(atoms.coeffs), _cabi.ptr(atoms.dw),
        atoms.model, _cabi.ptr(prefactor_d), _cabi.ptr(F), _cabi.ptr(I), int(hkl_int_max), _cabi.ptr(scratch))
    _cabi.check(rc, "ds_structure_factors")


def structure_factors(structure, g_indices, g_hkls_array, debye_waller_factors=None,
                      scattering_params="lobato", prefactor=None, dev=None, want_F=True, want_I=True):
    """K1 on device. Returns (F [n,2] float64 tensor or None, I [n] float64 tensor or None)."""
    dev = device(dev)
    atoms = AtomTable(structure, debye_waller_factors, scattering_params, dev)
    hkl = torch.as_tensor(np.ascontiguousarray(np.asarray(g_indices, float).reshape(-1, 3)), device=dev)
    gn = torch.as_tensor(np.ascontiguousarray(np.asarray(g_hkls_array, float).reshape(-1)), device=dev)
    n_g = hkl.shape[0]
    if gn.shape[0] != n_g:
        raise ValueError("g_indices and g_hkls_array must have the same length")
    pre_d = None
    if prefactor is not None and not np.isscalar(prefactor):
        pre_d = torch.as_tensor(np.array(np.broadcast_to(np.asarray(prefactor, float), (n_g,))), device=dev)
    F = torch.empty((n_g, 2), dtype=torch.float64, device=dev) if want_F else None
    I = torch.empty((n_g,), dtype=torch.float64, device=dev) if want_I else None
    launch_structure_factors(atoms, hkl, gn, pre_d, F, I)
    if I is not None and prefactor is not None and np.isscalar(prefactor) and prefactor != 1:
        I *= float(prefactor)
    return F, I


# ----------------------------------------------------------------------------------------------
# per-phase g table
# ----------------------------------------------------------------------------------------------
@dataclass
class GTable:
    hkl: np.ndarray          # [n,3] integer Miller indices (host)
    xyz_host: np.ndarray     # [n,3] float64 Cartesian g of the unrotated crystal (host)
    xyz: torch.Tensor        # [n,3] float64 device
    f32: torch.Tensor        # [n,4] float32 device (gx, gy, gz, |g|^2)
    I0: torch.Tensor         # [n]   float64 device, |F(g)|^2
    g_max: float
    # scan-line description of the table (None when it is not made of lattice lines)
    line_g0: torch.Tensor | None = None     # [n_lines, 4] float32 device
    line_start: torch.Tensor | None = None  # [n_lines + 1 (padded)] int32 device
    line_step: np.ndarray | None = None     # [3] float64 host
    n_lines: int = 0
    marked: bool = False                    # extinct rows carry |g|^2 = +inf in f32 (ds_pack_gtable)
    rows: np.ndarray | None = None          # rows of the enumerated g set this table keeps (None = all of them)

    @property
    def n(self):
        return self.xyz.shape[0]


def find_lines(xyz):
    """Describe a g table as lattice lines g0 + i * step of consecutive rows, if it is one.

    Both enumerations of the reference list Miller indices with l running fastest, so consecutive rows
    differ by +-c* except where a line ends (or where (000) was removed).  Returns (starts [n_lines + 1],
    step [3]) or None when fewer than half of the row-to-row differences are the common step."""
    n = xyz.shape[0]
    if n < 64:
        return None
    d = np.round(np.diff(xyz, axis=0), 9)
    uniq, inverse, counts = np.unique(d, axis=0, return_inverse=True, return_counts=True)
    best = int(np.argmax(counts))
    if counts[best] < 0.5 * (n - 1) or not np.any(uniq[best]):
        return None
    same = inverse.reshape(-1) == best
    starts = np.concatenate([[0], np.nonzero(~same)[0] + 1, [n]]).astype(np.int32)
    step = (xyz[1:][same] - xyz[:-1][same]).mean(axis=0)
    # every row must sit on its line to float32 accuracy
    L = np.repeat(np.arange(len(starts) - 1), np.diff(starts))
    i = np.arange(n) - starts[L]
    if np.abs(xyz[starts[L]] + i[:, None] * step - xyz).max() > 1e-7 * max(1.0, float(np.abs(xyz).max())):
        return None
    return starts, step


class GTablePlan:
    """Everything about a (structure, g set) that does not change between library builds: the enumerated
    hkl / Cartesian g and the atom table, uploaded once.  ``run()`` enqueues K1 + the table packing on the
    current stream and returns a fresh ``GTable`` -- no host arithmetic, no host<->device synchronisation."""

    def __init__(self, structure, hkl, xyz, debye_waller_factors, scattering_params, dev=None):
        dev = device(dev)
        self.atoms = AtomTable(structure, debye_waller_factors, scattering_params, dev)
        self._set_rows(hkl, xyz, dev)

    def _set_rows(self, hkl, xyz, dev):
        self._live = {}   # extinct_rel_cut -> plan over the rows that can pass the minimum-intensity cut
        self.rows = None  # set on such a sub-plan: its rows' indices in the enumerated g set
        self.hkl = np.ascontiguousarray(hkl)
        self.xyz_host = np.ascontiguousarray(np.asarray(xyz, float))
        gnorm = np.sqrt((self.xyz_host ** 2).sum(axis=1))
        self.g_max = float(gnorm.max()) if gnorm.size else 0.0
        self.hkl_d = torch.as_tensor(self.hkl.astype(float), device=dev)
        # integer Miller indices of bounded magnitude: K1 may factorise the phases (large cells)
        self.hkl_int_max = int(np.abs(self.hkl).max()) if self.hkl.size and np.issubdtype(self.hkl.dtype, np.integer) else 0
        self.sf_scratch = None
        if 0 < self.hkl_int_max <= 127 and self.atoms.n_atoms >= 32:
            nb = int(_cabi.lib().ds_structure_factors_scratch_bytes(self.atoms.n_atoms, self.hkl_int_max))
            self.sf_scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.gnorm_d = torch.as_tensor(gnorm, device=dev)
        self.xyz_d = torch.as_tensor(self.xyz_host, device=dev)
        self.lines = None
        found = find_lines(self.xyz_host)
        if found is not None:
            starts, step = found
            g0 = np.zeros((len(starts) - 1, 4), dtype=np.float32)
            g0[:, :3] = self.xyz_host[starts[:-1]]
            pad = (-len(starts)) % 4
            starts_p = np.concatenate([starts, np.full(pad, starts[-1], np.int32)]).astype(np.int32)
            self.lines = (torch.as_tensor(g0, device=dev), torch.as_tensor(starts_p, device=dev),
                          np.ascontiguousarray(step, dtype=np.float64), len(starts) - 1)

    def run(self, extinct_rel_cut=0.0, compact=True):
        """``extinct_rel_cut`` > 0 (and a table that ends with the direct beam): rows whose |F|^2 is at most that
        fraction of |F(000)|^2 can never pass the minimum-intensity cut (see ds_pack_gtable in the header for the
        argument) and are left out.  ``compact=True`` drops them from the plan -- one synchronising analysis pass
        per (plan, cut), after which K1 and K2 only ever see the live rows (13/16 of a diamond-cubic table is
        systematically absent); ``compact=False`` keeps the table and marks them for K2's cull instead."""
        if compact and extinct_rel_cut > 0.0 and self.hkl.shape[0] and not np.any(self.hkl[-1]):
            return self._live_plan(float(extinct_rel_cut))._run(0.0)
        return self._run(extinct_rel_cut)

    def _live_plan(self, rel):
        sub = self._live.get(rel)
        if sub is None:
            I0 = self._run(0.0).I0.cpu().numpy()
            live = ~(I0 <= rel * I0[-1])          # same rule as the marking; NaN rows stay
            if live.mean() > 0.75:
                # few extinct rows: keep the table (and its lattice-line structure) and mark them for the cull
                sub = _MarkedPlan(self, rel)
            else:
                sub = object.__new__(GTablePlan)
                sub.atoms = self.atoms
                sub._set_rows(self.hkl[live], self.xyz_host[live], self.xyz_d.device)
                sub.rows = np.nonzero(live)[0]
            self._live[rel] = sub
        return sub

    def _run(self, extinct_rel_cut=0.0):
        n = self.xyz_d.shape[0]
        dev = self.xyz_d.device
        I0 = torch.empty((n,), dtype=torch.float64, device=dev)
        f32 = torch.empty((n, 4), dtype=torch.float32, device=dev)
        mark = bool(n) and extinct_rel_cut > 0.0 and not np.any(self.hkl[-1])   # last row = (000)
        if n:
            launch_structure_factors(self.atoms, self.hkl_d, self.gnorm_d, None, None, I0,
                                     hkl_int_max=self.hkl_int_max if self.sf_scratch is not None else 0,
                                     scratch=self.sf_scratch)
            _cabi.check(_cabi.lib().ds_pack_gtable(_stream(), n, _cabi.ptr(self.xyz_d), _cabi.ptr(f32),
                                                   _cabi.ptr(I0) if mark else None, n - 1 if mark else -1,
                                                   float(extinct_rel_cut) if mark else 0.0), "ds_pack_gtable")
        gt = GTable(hkl=self.hkl, xyz_host=self.xyz_host, xyz=self.xyz_d, f32=f32, I0=I0, g_max=self.g_max,
                    marked=bool(mark), rows=self.rows)
        if self.lines is not None:
            gt.line_g0, gt.line_start, gt.line_step, gt.n_lines = self.lines
        return gt


class _MarkedPlan:
    """A plan whose extinct rows are marked at pack time instead of being dropped (see GTablePlan.run)."""

    def __init__(self, plan, rel):
        self.plan, self.rel = plan, rel

    def _run(self, _unused=0.0):
        return self.plan._run(self.rel)


def make_gtable(structure, hkl, xyz, debye_waller_factors, scattering_params, dev=None):
    return GTablePlan(structure, hkl, xyz, debye_waller_factors, scattering_params, dev).run()


# ----------------------------------------------------------------------------------------------
# K2
# ----------------------------------------------------------------------------------------------
@dataclass
class SpotTable:
    """Padded per-rotation reflection lists on the device (row r holds count[r] entries)."""
    count: torch.Tensor      # [n_rot] int32
    g_index: torch.Tensor    # [n_rot, cap] int32
    xyz: torch.Tensor        # [n_rot, cap, 3] float64
    intensity: torch.Tensor  # [n_rot, cap] float64
    exc: torch.Tensor | None
    cap: int
    # [1] int32 device word: the largest number of reflections any rotation produced BEFORE the minimum-intensity
    # cut; rows are only complete when it is <= cap (checked by simulate(check_overflow=True), else by the caller)
    max_count: torch.Tensor | None = None

    @property
    def n_rot(self):
        return self.count.shape[0]


def estimate_cap(n_g, g_max, s_max, precession_rad=0.0):
    thick = s_max + 0.7 * g_max * abs(precession_rad)
    mean = 1.5 * n_g * thick / max(g_max, 1e-6)
    cap = int(8 * mean + 32)
    cap = min(cap, n_g)
    return max(32, (cap + 31) // 32 * 32)


def simulate(gt: GTable, quats, wavelength, s_max, width, model, minima_number=5.0,
             precession_rad=0.0, min_intensity=1e-20, cap=None, want_exc=False, check_overflow=True):
    """Run K2 for active unit quaternions ``quats`` ([n,4] float64; numpy or device tensor)."""
    dev = gt.xyz.device
    if isinstance(quats, torch.Tensor):
        q = quats.to(device=dev, dtype=torch.float64).contiguous()
    else:
        q = torch.as_tensor(np.ascontiguousarray(np.asarray(quats, float).reshape(-1, 4)), device=dev)
    n_rot = q.shape[0]
    if cap is None:
        cap = estimate_cap(gt.n, gt.g_max, s_max, precession_rad)
    model_id = SHAPE_MODEL_IDS[model] if isinstance(model, str) else int(model)
    while True:
        count = torch.empty((n_rot,), dtype=torch.int32, device=dev)
        g_index = torch.empty((n_rot, cap), dtype=torch.int32, device=dev)
        xyz = torch.empty((n_rot, cap, 3), dtype=torch.float64, device=dev)
        inten = torch.empty((n_rot, cap), dtype=torch.float64, device=dev)
        exc = torch.empty((n_rot, cap), dtype=torch.float64, device=dev) if want_exc else None
        max_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        if gt.n == 0 or n_rot == 0:
            count.zero_()
            return SpotTable(count, g_index, xyz, inten, exc, cap)
        # measured (tools/bench_configs.py): once the extinct rows are marked, the plain cull over the packed table
        # beats the scan-line cull of the warp-per-rotation kernel, which rebuilds every row of a line
        # (DS_SIM_LINES=1 still forces it); large tables go to the CTA-per-rotation kernel, whose interval expansion
        # pays with or without marks (ds_simulate decides from the slab / step ratio)
        n_lines = gt.n_lines if (not gt.marked or gt.n >= 4096 or _cabi.get_option("sim_lines") == 1) else 0
        rc = _cabi.lib().ds_simulate(
            _stream(), n_rot, _cabi.ptr(q), gt.n, _cabi.ptr(gt.xyz), _cabi.ptr(gt.f32), _cabi.ptr(gt.I0),
            float(gt.g_max), 1.0 / float(wavelength), float(s_max), float(width), model_id,
            float(minima_number), float(precession_rad), float(min_intensity), cap,
            _cabi.ptr(count), _cabi.ptr(g_index), _cabi.ptr(xyz), _cabi.ptr(inten), _cabi.ptr(exc),
            _cabi.ptr(max_count), int(n_lines), _cabi.ptr(gt.line_g0), _cabi.ptr(gt.line_start),
            None if gt.line_step is None else gt.line_step.ctypes.data_as(_cabi.c_void_p))
        _cabi.check(rc, "ds_simulate")
        if not check_overflow:
            return SpotTable(count, g_index, xyz, inten, exc, cap, max_count)
        need = int(max_count.item())
        if need <= cap:
            return SpotTable(count, g_index, xyz, inten, exc, cap, max_count)
        cap = (need + 31) // 32 * 32


# ----------------------------------------------------------------------------------------------
# K3
# ----------------------------------------------------------------------------------------------
_SCRATCH = {}


def _render_scratch(dev, n, cap):
    """Device scratch of ds_render (ticket words + prepared template records), one buffer per (device, stream),
    grown on demand.  Streams under CUDA-graph capture get their own allocation from the graph's pool."""
    need = int(_cabi.lib().ds_render_scratch_bytes(int(n), int(cap)))
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(need, dtype=torch.uint8, device=dev)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < need:
        t = _SCRATCH[key] = torch.empty(max(need, 4096), dtype=torch.uint8, device=dev)
    return t


def quantize_u16(images, out=None):
    """Normalised float32 templates -> uint16 (rint(v * 65535)) on the device: the optional half-size export of a
    template library (quantisation 1.5e-5 of the peak).  ``out``: a uint16 device tensor of the same shape."""
    if images.dtype != torch.float32 or not images.is_contiguous():
        raise ValueError("quantize_u16 needs a contiguous float32 device tensor")
    if out is None:
        out = torch.empty(images.shape, dtype=torch.uint16, device=images.device)
    elif out.dtype != torch.uint16 or out.shape != images.shape or not out.is_contiguous():
        raise ValueError("quantize_u16: `out` must be a contiguous uint16 tensor of the same shape")
    _cabi.check(_cabi.lib().ds_quantize_u16(_stream(), images.numel(), _cabi.ptr(images), _cabi.ptr(out)), "ds_quantize_u16")
    return out


def render_launch_count(cap, shape, sigma, fast=True, mean_spots=None):
    """How many kernels ds_render launches for this configuration (1, or 2 when the tcgen05 path with its prepare
    pass is taken) -- the library's own dispatch rule, for honest launch accounting."""
    return int(_cabi.lib().ds_render_launch_count(int(cap), int(shape[0]), int(shape[1]), gaussian_radius(sigma),
                                                  2 if fast == "bare" else int(bool(fast)),
                                                  0.0 if mean_spots is None else float(mean_spots)))


def gaussian_radius(sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d: lw = int(truncate * sd + 0.5)."""
    return int(truncate * float(sigma) + 0.5)


def render(count, xyz, intensity, shape, sigma, calibration, center, in_plane_angle=0.0, mirrored=False,
           fast=True, normalize=True, clip_threshold=1.0, out=None, mean_spots=None):
    """Run K3. ``count`` [n] int32, ``xyz`` [n,cap,3] f64, ``intensity`` [n,cap] f64 device tensors.
    ``fast``: True = integer-pixel branch, False = sub-pixel branch on the in-frame spots, "bare" = sub-pixel
    branch without the in-frame selection (detector_functions.get_pattern_from_pixel_coordinates_and_intensities).
    ``mean_spots``: the mean of ``count`` if the caller knows it without synchronising (a schedule hint only)."""
    dev = xyz.device
    n, cap = intensity.shape
    H, W = int(shape[0]), int(shape[1])
    if out is None:
        out = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    else:
        assert out.shape == (n, H, W) and out.dtype == torch.float32 and out.is_contiguous()
    rc = _cabi.lib().ds_render(
        _stream(), n, cap, _cabi.ptr(count), _cabi.ptr(xyz), _cabi.ptr(intensity), H, W,
        float(calibration), float(center[0]), float(center[1]), float(in_plane_angle), int(bool(mirrored)),
        (2 if fast == "bare" else int(bool(fast))), float(sigma), gaussian_radius(sigma), float(clip_threshold), int(bool(normalize)),
        _cabi.ptr(out), _cabi.ptr(_render_scratch(dev, n, cap)), 0.0 if mean_spots is None else float(mean_spots))
    _cabi.check(rc, "ds_render")
    return out


# ----------------------------------------------------------------------------------------------
# polar flattening
# ----------------------------------------------------------------------------------------------
def polar_flatten(count, xyz, intensity, max_spots, radial_axes=None, azimuthal_axes=None):
    """Device polar flattening of packed spot rows: returns (r, theta, intensity) float64 tensors
    [n, max_spots] (r / theta hold axis indices when both axes are given)."""
    dev = xyz.device
    n, cap = intensity.shape
    r = torch.empty((n, max_spots), dtype=torch.float64, device=dev)
    t = torch.empty_like(r)
    i = torch.empty_like(r)
    rad = az = None
    if radial_axes is not None and azimuthal_axes is not None:
        rad = torch.as_tensor(np.ascontiguousarray(np.asarray(radial_axes, float)), device=dev)
        az = torch.as_tensor(np.ascontiguousarray(np.asarray(azimuthal_axes, float)), device=dev)
    rc = _cabi.lib().ds_polar_flatten(
        _stream(), n, cap, _cabi.ptr(count), _cabi.ptr(xyz), _cabi.ptr(intensity), int(max_spots),
        0 if rad is None else rad.numel(), _cabi.ptr(rad), 0 if az is None else az.numel(), _cabi.ptr(az),
        _cabi.ptr(r), _cabi.ptr(t), _cabi.ptr(i))
    _cabi.check(rc, "ds_polar_flatten")
    return r, t, i


def library_pixel_coords(count, xyz, calibration, half_shape, offset=(0.0, 0.0)):
    """rint((xy + offset) / calibration + half_shape) for every stored reflection: int32 [n, cap, 2]."""
    dev = xyz.device
    n, cap = xyz.shape[0], xyz.shape[1]
    cal = np.broadcast_to(np.asarray(calibration, float), (2,))
    half = np.broadcast_to(np.asarray(half_shape, float), (2,))
    out = torch.empty((n, cap, 2), dtype=torch.int32, device=dev)
    rc = _cabi.lib().ds_library_pixel_coords(_stream(), n, cap, _cabi.ptr(count), _cabi.ptr(xyz), float(cal[0]),
                                             float(cal[1]), float(offset[0]), float(offset[1]), float(half[0]),
                                             float(half[1]), _cabi.ptr(out))
    _cabi.check(rc, "ds_library_pixel_coords")
    return out
