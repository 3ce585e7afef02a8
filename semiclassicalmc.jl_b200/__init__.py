"""semiclassicalmc.jl_b200 -- B200-native Metropolis / simulated-annealing sweep for su(4) product-state spin-valley
models, a drop-in for that path of lgresista/SemiClassicalMC.jl (imported as `scmc_b200` through the shim at the repo root).

Host-side mirror of the reference's Julia API (src/SemiClassicalMC.jl:19-87) on top of the C ABI in include/scmc.h;
Julia's `f!` names are spelled `f_` here.
"""
from . import _lib
from ._lib import ScmcError      # the exception every failing library call raises
from .build import build as build_library


def __getattr__(name):   # lazy: the API modules need numpy only, the CUDA library is loaded on first use
    import importlib
    _mods = ("generators", "lattice", "configuration", "models", "observables", "montecarlo", "launcher", "binning", "distributed", "io", "hdf5")
    if name in _mods:
        return importlib.import_module(f".{name}", __name__)
    for mod in ("generators", "lattice", "configuration", "models", "observables", "montecarlo", "launcher", "binning", "distributed", "io", "hdf5"):
        try:
            m = importlib.import_module(f".{mod}", __name__)
        except ModuleNotFoundError:
            continue
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(name)
