"""Build the C-ABI shared library ``libdiffsims_b200.so`` in-tree with nvcc for sm_100a.

    python -m diffsims_b200.build

nvcc cross-compiles without a GPU.  The library links the CUDA runtime statically and has
no torch dependency; it shares the primary context with whatever created the device
pointers it is handed.
"""
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libdiffsims_b200.so"
SOURCES = ["cabi.cu", "structure_factor.cu", "simulate.cu", "render.cu", "render_pipe.cu", "render_umma.cu", "render_prep.cu", "render_rows.cu", "polar.cu", "beam_grid.cu", "so3_grid.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "diffsims_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """``defines`` / ``out``: instrumented variants for the tools (e.g. DS_PROF -> per-role cycle counters of the
    tcgen05 render kernel); the product library is always built without defines."""
    out = Path(out) if out else LIB
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(out), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
