"""``SimulationGenerator`` -- kinematical diffraction simulation, B200-native.

Drop-in mirror of diffsims/generators/simulation_generator.py:63-259 (constructor, ``wavelength``,
``calculate_diffraction2d``).  The reference loops over rotations in Python; here one phase is
  1. enumerated on the host (integer hkl inside the reciprocal sphere, reference order),
  2. given structure factors by the K1 kernel once (the reference recomputes them per rotation),
  3. simulated for ALL rotations by one launch of the fused K2 kernel,
and the result stays packed on the device inside ``Simulation2D`` (objects are built lazily).
"""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np
import torch
from tqdm import tqdm

from .. import engine
from ..crystal import Rotation
from ..crystallography import DiffractingVector, g_set_from_min_dspacing
from ..simulations.simulation2d import PackedVectors, Simulation2D, _is_phase, _is_rotation
from ..utils import shape_factor_models as sfm
from ..utils.sim_utils import get_electron_wavelength

__all__ = ["SimulationGenerator"]

_shape_factor_model_mapping = {
    "linear": sfm.linear,
    "atanc": sfm.atanc,
    "sinc": sfm.sinc,
    "sin2c": sfm.sin2c,
    "lorentzian": sfm.lorentzian,
}


class SimulationGenerator:
    """A class for generating kinematic diffraction simulations."""

    def __repr__(self):
        return (f"SimulationGenerator(accelerating_voltage={self.accelerating_voltage}, "
                f"scattering_params={self.scattering_params}, "
                f"approximate_precession={self.approximate_precession})")

    def __init__(self, accelerating_voltage: float = 200, scattering_params: str = "lobato",
                 precession_angle: float = 0, shape_factor_model: str = "lorentzian",
                 approximate_precession: bool = True, minimum_intensity: float = 1e-20, **kwargs):
        self.accelerating_voltage = accelerating_voltage
        self.precession_angle = np.abs(precession_angle)
        self.approximate_precession = approximate_precession
        if isinstance(shape_factor_model, str):
            if shape_factor_model in _shape_factor_model_mapping.keys():
                self.shape_factor_model = _shape_factor_model_mapping[shape_factor_model]
            else:
                raise NotImplementedError(
                    f"{shape_factor_model} is not a recognized shape factor "
                    f"model, choose from: {_shape_factor_model_mapping.keys()} "
                    f"or provide your own function.")
        else:
            self.shape_factor_model = shape_factor_model
        self.minimum_intensity = minimum_intensity
        self.shape_factor_kwargs = kwargs
        if scattering_params in ["lobato", "xtables", None]:
            self.scattering_params = scattering_params
        else:
            raise NotImplementedError(
                "The scattering parameters `{}` is not implemented. "
                "See documentation for available "
                "implementations.".format(scattering_params))

    @property
    def wavelength(self):
        return get_electron_wavelength(self.accelerating_voltage)

    # ------------------------------------------------------------------------------------------
    def _native_model(self):
        """(native model name or None for a Python callable, minima_number)."""
        minima = float(self.shape_factor_kwargs.get("minima_number", 5))
        if self.precession_angle != 0 and self.approximate_precession:
            return "lorentzian_precession", minima   # shape_factor_model is ignored, as in the reference
        # precession_angle != 0 and not approximate: K2 averages the native model over the precession
        # circle (_shape_factor_precession); a Python callable cannot be integrated in the kernel
        name = sfm.NATIVE.get(self.shape_factor_model)
        if name is not None and set(self.shape_factor_kwargs) <= {"minima_number"}:
            return name, minima
        if self.precession_angle != 0:
            raise NotImplementedError(
                "approximate_precession=False needs a native shape factor model "
                "(binary, linear, sinc, sin2c, atanc, lorentzian)")
        return None, minima

    def _g_plan(self, phase, reciprocal_radius, with_direct_beam, debye_waller_factors):
        """Host enumeration (from_min_dspacing, reciprocal_lattice_vector.py:1077-1142) + uploads."""
        lat = phase.structure.lattice
        hkl = g_set_from_min_dspacing(lat, 1 / reciprocal_radius, include_zero_vector=with_direct_beam)
        if with_direct_beam:
            # get_intersecting_reflections stacks ANOTHER (000) onto the rotated set
            # (simulation_generator.py:351-353): the direct beam is listed twice, as in the reference
            hkl = np.vstack([hkl, np.zeros((1, 3), dtype=hkl.dtype)])
        xyz = hkl.astype(float) @ np.asarray(lat.recbase, dtype=float).T
        return engine.GTablePlan(phase.structure, hkl, xyz, debye_waller_factors, self.scattering_params)

    def _extinct_rel_cut(self, with_direct_beam):
        """Relative |F|^2 below which a reflection can never pass ``minimum_intensity`` (0 = do not mark): needs
        the direct beam in the set (it is always excited, so max(I) >= sf(0) |F(000)|^2) and a shape factor
        bounded by its value at s = 0 -- see ds_pack_gtable in include/diffsims_b200.h."""
        model, _ = self._native_model()
        if (with_direct_beam and self.precession_angle == 0 and self.minimum_intensity > 0
                and model in ("binary", "linear", "atanc", "lorentzian")):
            return min(0.5 * self.minimum_intensity, 0.5)
        return 0.0

    def _g_table(self, phase, reciprocal_radius, with_direct_beam, debye_waller_factors):
        """Per-phase g table with structure factors (K1)."""
        plan = self._g_plan(phase, reciprocal_radius, with_direct_beam, debye_waller_factors)
        return plan.run(self._extinct_rel_cut(with_direct_beam))

    def _simulate_phase(self, phase, rotation, reciprocal_radius, with_direct_beam, max_excitation_error,
                        shape_factor_width, debye_waller_factors):
        gt = self._g_table(phase, reciprocal_radius, with_direct_beam, debye_waller_factors)
        if shape_factor_width is None:
            shape_factor_width = max_excitation_error
        # vecs = ~rotation * g (_diffracting_vector.py:160): the kernel applies the ACTIVE matrix of the
        # quaternion it is given, so hand it the inverse rotation
        quats = np.asarray((~rotation).data, dtype=float).reshape(-1, 4)
        model, minima = self._native_model()
        prec = float(np.deg2rad(self.precession_angle))
        if model is not None:
            spots = engine.simulate(gt, quats, self.wavelength, max_excitation_error, shape_factor_width,
                                    model, minima, prec, self.minimum_intensity)
        else:
            spots = self._apply_callable(gt, quats, max_excitation_error, shape_factor_width, prec)
        return PackedVectors(phase, rotation, spots, gt.hkl)

    def _apply_callable(self, gt, quats, s_max, width, prec):
        """A Python ``shape_factor_model`` cannot run inside the kernel: K2 returns the reflections that
        pass the excitation-error cut with their excitation errors and |F|^2; the callable, the product
        and the minimum_intensity cut (simulation_generator.py:387-394, :237) are applied to the packed
        arrays here."""
        raw = engine.simulate(gt, quats, self.wavelength, s_max, width, "return_s", 5.0, prec, 0.0,
                              want_exc=True)
        count = raw.count.cpu().numpy()
        s = raw.exc.cpu().numpy()
        valid = np.arange(raw.cap)[None, :] < count[:, None]
        shape = np.zeros_like(s)
        shape[valid] = np.asarray(
            self.shape_factor_model(s[valid], width, **self.shape_factor_kwargs), dtype=float)
        inten = shape * raw.intensity.cpu().numpy()
        inten[~valid] = -np.inf
        mx = np.where(count > 0, inten.max(axis=1, initial=-np.inf), 0.0)
        keep = valid & (inten > (mx * self.minimum_intensity)[:, None])
        order = np.argsort(~keep, axis=1, kind="stable")  # kept entries first, original order
        dev = raw.count.device
        o = torch.as_tensor(order, device=dev)
        inten[~keep] = 0.0
        return engine.SpotTable(
            count=torch.as_tensor(keep.sum(axis=1).astype(np.int32), device=dev),
            g_index=torch.gather(raw.g_index, 1, o.to(torch.int64)).contiguous(),
            xyz=torch.gather(raw.xyz, 1, o.to(torch.int64)[:, :, None].expand(-1, -1, 3)).contiguous(),
            intensity=torch.gather(torch.as_tensor(inten, device=dev), 1, o.to(torch.int64)).contiguous(),
            exc=torch.gather(raw.exc, 1, o.to(torch.int64)).contiguous(), cap=raw.cap)

    # ------------------------------------------------------------------------------------------
    def calculate_diffraction2d(self, phase, rotation=None, reciprocal_radius: float = 1.0,
                                with_direct_beam: bool = True, max_excitation_error: float = 1e-2,
                                shape_factor_width: float = None, debye_waller_factors: dict = None,
                                show_progressbar: bool = False):
        """Calculates the diffraction pattern for one or more phases given a list of rotations for each
        phase (simulation_generator.py:134-259).  Same parameters and return type as the reference."""
        if rotation is None:
            rotation = Rotation.from_euler((0, 0, 0), degrees=True)
        if _is_phase(phase):
            phase = [phase]
        if _is_rotation(rotation):
            rotation = [rotation]
        if len(phase) != len(rotation):
            raise ValueError("The number of phases and rotations must be equal. "
                             f"Got {len(phase)} phases and {len(rotation)} rotations.")
        if debye_waller_factors is None:
            debye_waller_factors = {}

        vectors = []
        for p, rotate in zip(phase, rotation):
            bar = tqdm(desc=p.name, total=rotate.size) if show_progressbar else None
            packed = self._simulate_phase(p, rotate, reciprocal_radius, with_direct_beam,
                                          max_excitation_error, shape_factor_width, debye_waller_factors)
            if bar is not None:
                torch.cuda.current_stream().synchronize()
                bar.update(rotate.size)
                bar.close()
            vectors.append(packed)

        if len(phase) == 1:
            vectors = vectors[0]
            phase = phase[0]
            rotation = rotation[0]
            if rotation.size == 1:
                vectors = vectors[0]

        return Simulation2D(phases=phase, coordinates=vectors, rotations=rotation,
                            simulation_generator=self, reciprocal_radius=reciprocal_radius)

    def get_intersecting_reflections(self, recip, rot, wavelength, max_excitation_error,
                                     shape_factor_width=None, with_direct_beam=True):
        """Reflections of ``recip`` that intersect the Ewald sphere for ONE rotation
        (simulation_generator.py:319-412): returns (DiffractingVector, hkl, shape_factor).

        Runs K2 on the given vector set with unit structure factors so that the returned intensity
        is the shape factor itself."""
        if rot.size != 1:
            raise ValueError("Rotation must be a single rotation")
        xyz = np.asarray(recip.data, dtype=float).reshape(-1, 3)
        hkl = np.asarray(recip.hkl, dtype=float).reshape(-1, 3)
        if with_direct_beam:
            xyz = np.vstack([xyz, [0, 0, 0]])
            hkl = np.vstack([hkl, [0, 0, 0]])
        dev = engine.device()
        xyz_d = torch.as_tensor(np.ascontiguousarray(xyz), device=dev)
        f32 = torch.empty((xyz.shape[0], 4), dtype=torch.float32, device=dev)
        engine._cabi.check(engine._cabi.lib().ds_pack_gtable(
            engine._stream(), xyz.shape[0], engine._cabi.ptr(xyz_d), engine._cabi.ptr(f32), None, -1, 0.0),
            "ds_pack_gtable")
        gt = engine.GTable(hkl=hkl, xyz_host=xyz, xyz=xyz_d, f32=f32,
                           I0=torch.ones(xyz.shape[0], dtype=torch.float64, device=dev),
                           g_max=float(np.sqrt((xyz ** 2).sum(axis=1)).max()) if xyz.size else 0.0)
        if shape_factor_width is None:
            shape_factor_width = max_excitation_error
        model, minima = self._native_model()
        prec = float(np.deg2rad(self.precession_angle))
        quats = np.asarray((~rot).data, dtype=float).reshape(-1, 4)
        raw = engine.simulate(gt, quats, wavelength, max_excitation_error, shape_factor_width,
                              model if model is not None else "return_s", minima, prec, -1.0, want_exc=True)
        n = int(raw.count[0])
        idx = raw.g_index[0, :n].cpu().numpy()
        if model is not None:
            shape_factor = raw.intensity[0, :n].cpu().numpy()
        else:
            shape_factor = self.shape_factor_model(raw.exc[0, :n].cpu().numpy(), shape_factor_width,
                                                   **self.shape_factor_kwargs)
        G = np.asarray(rot.to_matrix()).reshape(3, 3)

        def rotated_phase():
            ph = recip.phase.deepcopy()
            ph.structure.lattice.setLatPar(baserot=np.asarray(ph.structure.lattice.baserot) @ G)
            return ph

        dv = DiffractingVector(rotated_phase, xyz=raw.xyz[0, :n].cpu().numpy())
        return dv, hkl[idx], shape_factor

    def calculate_diffraction1d(self, *args, **kwargs):  # pragma: no cover
        raise NotImplementedError("the 1-D powder profile is outside the template-simulation path")
