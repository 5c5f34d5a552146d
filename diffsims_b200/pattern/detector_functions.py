"""The rasteriser of the template path behind the reference's function name.

``get_pattern_from_pixel_coordinates_and_intensities`` (diffsims/pattern/detector_functions.py:251-311) is what
``Simulation2D.get_diffraction_pattern`` and ``DiffractionSimulation.get_diffraction_pattern`` hand their pixel
coordinates to.  Here it is one ``ds_render`` launch (calibration 1, no rotation, origin at pixel 0): integer
coordinates take the assignment + Gaussian-filter branch, float coordinates the sub-pixel branch.  The detector
noise models of the same reference module are not on the simulation path and are not provided.
"""
import numpy as np
import torch

from .. import engine

__all__ = ["get_pattern_from_pixel_coordinates_and_intensities"]


def get_pattern_from_pixel_coordinates_and_intensities(coordinates, intensities, shape, sigma, clip_threshold=1):
    """Diffraction pattern [H, W] (float64) from spot pixel coordinates (n, 2) or (n, 3), x first, and
    intensities (n,).  Same arguments and branches as the reference (:251-311)."""
    coordinates = np.asarray(coordinates)
    intensities = np.asarray(intensities, dtype=np.float64).reshape(-1)
    H, W = int(shape[0]), int(shape[1])
    integer = np.issubdtype(coordinates.dtype, np.integer)
    xy = np.array(coordinates[:, :2], dtype=np.int64 if integer else np.float64)
    if xy.shape[0] != intensities.shape[0]:
        raise ValueError(f"{xy.shape[0]} coordinates but {intensities.shape[0]} intensities")
    if integer:
        # numpy fancy-index assignment out[y, x] = I (:297): negative indices wrap, others raise
        for axis, n in ((0, W), (1, H)):
            if xy.shape[0] and (xy[:, axis].min() < -n or xy[:, axis].max() >= n):
                raise IndexError(f"index out of bounds for axis {1 - axis} with size {n}")
            xy[:, axis] = np.where(xy[:, axis] < 0, xy[:, axis] + n, xy[:, axis])
    n = xy.shape[0]
    dev = engine.device()
    cap = max(32, (n + 31) // 32 * 32)
    xyz = np.zeros((1, cap, 3))
    xyz[0, :n, :2] = xy
    inten = np.zeros((1, cap))
    inten[0, :n] = intensities
    out = engine.render(torch.tensor([n], dtype=torch.int32, device=dev), torch.as_tensor(xyz, device=dev),
                        torch.as_tensor(inten, device=dev), (H, W), sigma, 1.0, (0.0, 0.0),
                        fast=True if integer else "bare", normalize=False, clip_threshold=clip_threshold)
    return out[0].cpu().numpy().astype(np.float64)
