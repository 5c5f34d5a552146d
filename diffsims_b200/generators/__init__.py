from .diffraction_generator import DiffractionGenerator
from .library_generator import DiffractionLibraryGenerator
from .simulation_generator import SimulationGenerator

__all__ = ["DiffractionGenerator", "DiffractionLibraryGenerator", "SimulationGenerator"]
