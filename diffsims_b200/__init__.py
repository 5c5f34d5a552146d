"""diffsims_b200 -- B200-native kinematical template simulation behind the diffsims API.

Only the hot path named in BASELINE.json is implemented: structure factors (K1),
the fused per-rotation simulate kernel (K2) and the rasteriser (K3), reached through the C ABI in
include/diffsims_b200.h.  Importing the package does not need a GPU; calling it does, and there is no
CPU fallback.
"""
__version__ = "0.1.0"

from .crystal import Atom, Lattice, Phase, Rotation, Structure  # noqa: F401
from .crystallography import DiffractingVector, ReciprocalLatticeVector  # noqa: F401
from .generators import DiffractionGenerator, DiffractionLibraryGenerator, SimulationGenerator  # noqa: F401
from .libraries import DiffractionLibrary, StructureLibrary  # noqa: F401
from .simulations import Simulation2D  # noqa: F401
from .sims import DiffractionSimulation  # noqa: F401
