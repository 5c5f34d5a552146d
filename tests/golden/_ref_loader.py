"""Load the importable LEAF modules of the reference by file path (dev container only).

The reference package itself cannot be imported here (orix, diffpy.structure,
transforms3d and matplotlib are not installed -- SURVEY.md §8c).  Its leaf modules
on the hot path only need numpy/scipy/numba plus a few names from those
packages, so this loader registers empty placeholder packages and executes the
reference files *where they lie* under /root/reference under their canonical
module names.  Nothing is copied.

Placeholders that replace third-party code (everything else is the reference's
own code, unmodified):

* ``diffpy.structure``          -> empty module (sim_utils only needs the import to succeed;
                                   the structures passed in are the duck-typed stand-ins
                                   from ``diffsims_b200.crystal``)
* ``transforms3d.euler.euler2mat(ai, aj, ak, axes='rzxz')`` -> Rz(ai) Rx(aj) Rz(ak)
  (the published definition of rotating-frame zxz); used only by the OLD api
  ``calculate_ed_data``
* ``matplotlib.pyplot``, ``PIL``, ``diffsims.utils.fourier_transform`` -> empty (plot/FFT
  helpers that the hot path never calls)
* ``orix.sampling.sample_generators``, ``orix.quaternion.rotation`` -> empty names (imported by
  rotation_list_generators.py but not used by ``get_beam_directions_grid``)

Used only by tests/golden/make_golden.py.
"""
import importlib.util
import math
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")


def _pkg(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return sys.modules[name]


def _load(name, rel):
    if name in sys.modules and getattr(sys.modules[name], "__file__", None):
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_pkg(parent), child, mod)
    return mod


def _euler2mat(ai, aj, ak, axes="rzxz"):
    assert axes == "rzxz"

    def rz(t):
        c, s = math.cos(t), math.sin(t)
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])

    def rx(t):
        c, s = math.cos(t), math.sin(t)
        return np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]])
    return rz(ai) @ rx(aj) @ rz(ak)


def load_reference():
    """Returns a namespace with the reference leaf modules."""
    for name in ("diffsims", "diffsims.utils", "diffsims.pattern", "diffsims.structure_factor",
                 "diffsims.generators", "diffsims.sims", "diffsims.libraries",
                 "diffpy", "transforms3d", "matplotlib", "PIL"):
        _pkg(name)
    sys.modules["diffpy.structure"] = types.ModuleType("diffpy.structure")
    sys.modules["diffpy"].structure = sys.modules["diffpy.structure"]
    eul = types.ModuleType("transforms3d.euler")
    eul.euler2mat = _euler2mat
    sys.modules["transforms3d.euler"] = eul
    sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for n in ("PIL.Image", "PIL.ImageDraw"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    sys.modules["PIL"].ImageDraw = sys.modules["PIL.ImageDraw"]
    ft = types.ModuleType("diffsims.utils.fourier_transform")
    ft.from_recip = None
    sys.modules["diffsims.utils.fourier_transform"] = ft

    ns = types.SimpleNamespace()
    _load("diffsims.utils.atomic_scattering_params", "diffsims/utils/atomic_scattering_params.py")
    _load("diffsims.utils.lobato_scattering_params", "diffsims/utils/lobato_scattering_params.py")
    _load("diffsims.structure_factor.atomic_scattering_parameters",
          "diffsims/structure_factor/atomic_scattering_parameters.py")
    ns.shape_factor_models = _load("diffsims.utils.shape_factor_models",
                                   "diffsims/utils/shape_factor_models.py")
    ns.detector_functions = _load("diffsims.pattern.detector_functions",
                                  "diffsims/pattern/detector_functions.py")
    ns.sim_utils = _load("diffsims.utils.sim_utils", "diffsims/utils/sim_utils.py")
    ns.vector_utils = _load("diffsims.utils.vector_utils", "diffsims/utils/vector_utils.py")
    _load("diffsims.utils.mask_utils", "diffsims/utils/mask_utils.py")
    ns.diffraction_simulation = _load("diffsims.sims.diffraction_simulation",
                                      "diffsims/sims/diffraction_simulation.py")
    ns.diffraction_generator = _load("diffsims.generators.diffraction_generator",
                                     "diffsims/generators/diffraction_generator.py")
    ns.sphere_mesh_generators = _load("diffsims.generators.sphere_mesh_generators",
                                      "diffsims/generators/sphere_mesh_generators.py")
    # rotation_list_generators imports two orix names at module level that get_beam_directions_grid never
    # calls (get_sample_fundamental / get_sample_local / Rotation): empty placeholders
    for n in ("orix", "orix.sampling", "orix.quaternion"):
        _pkg(n)
    sg = types.ModuleType("orix.sampling.sample_generators")
    sg.get_sample_fundamental = sg.get_sample_local = None
    sys.modules["orix.sampling.sample_generators"] = sg
    rq = types.ModuleType("orix.quaternion.rotation")
    rq.Rotation = None
    sys.modules["orix.quaternion.rotation"] = rq
    ns.rotation_list_generators = _load("diffsims.generators.rotation_list_generators",
                                        "diffsims/generators/rotation_list_generators.py")
    return ns
