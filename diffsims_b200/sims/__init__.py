from .diffraction_simulation import DiffractionSimulation

__all__ = ["DiffractionSimulation"]
