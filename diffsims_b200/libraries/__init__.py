from .diffraction_library import DiffractionLibrary, load_DiffractionLibrary
from .structure_library import StructureLibrary

__all__ = ["DiffractionLibrary", "StructureLibrary", "load_DiffractionLibrary"]
