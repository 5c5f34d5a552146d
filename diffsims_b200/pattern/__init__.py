"""Pattern rasterisation (the part of diffsims.pattern that the simulation path uses)."""
from . import detector_functions  # noqa: F401
