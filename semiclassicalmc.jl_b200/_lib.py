"""ctypes binding of libscmc_b200.so -- 1:1 with include/scmc.h.  No fallback: if the CUDA library is missing or a
call fails, an exception is raised (the product never routes through a CPU path)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCMC_LIB") or os.path.join(HERE, "libscmc_b200.so")   # SCMC_LIB: alternative build (kernel experiments)
NOBS = 21

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_bp = C.POINTER(C.c_uint8)


class ScmcError(RuntimeError):
    pass


class RunParams(C.Structure):
    _fields_ = [("n_therm", C.c_int64), ("n_measure", C.c_int64), ("measurement_rate", C.c_int64), ("hits", C.c_int32),
                ("stop_sweep", C.c_int32), ("seed", C.c_uint64), ("sweep0", C.c_uint64), ("T", c_dp), ("T_i", c_dp), ("sigma", c_dp)]


class RunResults(C.Structure):
    _fields_ = [("series", c_dp), ("n_rows", C.c_int64), ("mean", c_dp), ("acc_rate", c_dp), ("therm_energy", c_dp)]


class AnnealParams(C.Structure):
    _fields_ = [("T_i", C.c_double), ("T_fac", C.c_double), ("N_max", C.c_int64), ("N_per_T", C.c_int64),
                ("min_acc_per_site", C.c_int64), ("min_accrate", C.c_double), ("sigma_min", C.c_double), ("sigma0", C.c_double),
                ("hits", C.c_int32), ("log_rate", C.c_int32), ("seed", C.c_uint64)]


class AnnealResults(C.Structure):
    _fields_ = [("E", c_dp), ("beta", c_dp), ("sigma", c_dp), ("sweeps", c_lp), ("energy_log", c_dp), ("beta_log", c_dp), ("n_log_cap", C.c_int64),
                ("n_log", C.c_int64)]


# symbol -> argtypes; every symbol of include/scmc.h (checked by tests/test_abi.py)
SIGNATURES = {
    "scmc_last_error": (C.c_char_p, []),
    "scmc_version": (C.c_int32, []),
    "scmc_device_count": (C.c_int32, [c_ip]),
    "scmc_create": (C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                c_ip, c_ip, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "scmc_destroy": (C.c_int32, [C.c_void_p]),
    "scmc_set_stream": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "scmc_info": (C.c_int32, [C.c_void_p, c_lp]),
    "scmc_verify_tables": (C.c_int32, [C.c_void_p, c_lp]),
    "scmc_set_state": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, c_dp, c_dp]),
    "scmc_get_state": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, c_dp, c_dp]),
    "scmc_randomize": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "scmc_get_T": (C.c_int32, [C.c_void_p, C.c_int64, c_dp, c_dp]),
    "scmc_energy": (C.c_int32, [C.c_void_p, c_dp]),
    "scmc_local_field": (C.c_int32, [C.c_void_p, C.c_int64, c_dp]),
    "scmc_update_deterministic": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, c_dp, C.c_double, C.c_double, C.c_double, c_dp, c_ip]),
    "scmc_sweep_deterministic": (C.c_int32, [C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_bp, c_dp]),
    "scmc_sweep": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, c_dp, c_dp, C.c_uint64, C.c_uint64, c_lp, c_lp]),
    "scmc_sweep_async": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64]),
    "scmc_set_sweep_mode": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64]),
    "scmc_set_stream_version": (C.c_int32, [C.c_void_p, C.c_int32]),
    "scmc_set_lanes": (C.c_int32, [C.c_void_p, C.c_int32]),
    "scmc_measure": (C.c_int32, [C.c_void_p, c_dp]),
    "scmc_structure_factor": (C.c_int32, [C.c_void_p, c_dp, c_dp, C.c_int32, C.c_int32, C.c_double, c_dp]),
    "scmc_structure_factor_grid": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_ip, c_ip, C.c_int32, c_ip, c_dp, C.c_double, C.c_uint32, c_dp]),
    "scmc_structure_factor_accumulate": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_ip, c_ip, C.c_int32, c_ip, c_dp, C.c_double, C.c_uint32]),
    "scmc_structure_factor_fetch": (C.c_int32, [C.c_void_p, c_dp, C.POINTER(C.c_int64), C.c_int32]),
    "scmc_correlations": (C.c_int32, [C.c_void_p, C.c_int64, c_ip, C.c_int64, c_dp]),
    "scmc_sublattice_magnetization": (C.c_int32, [C.c_void_p, c_ip, C.c_int32, c_dp]),
    "scmc_chirality": (C.c_int32, [C.c_void_p, c_ip, c_dp]),
    "scmc_run_metropolis": (C.c_int32, [C.c_void_p, C.POINTER(RunParams), C.POINTER(RunResults)]),
    "scmc_anneal": (C.c_int32, [C.c_void_p, C.POINTER(AnnealParams), C.POINTER(AnnealResults)]),
    "scmc_optimize": (C.c_int32, [C.c_void_p, C.c_int64, c_dp]),
    # several GPUs of one box behind one object (single process, ncclAllGather of the measure! rows)
    "scmc_multi_create": (C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, c_ip, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                      c_ip, c_ip, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "scmc_multi_destroy": (C.c_int32, [C.c_void_p]),
    "scmc_multi_count": (C.c_int32, [C.c_void_p, c_ip, c_lp]),
    "scmc_multi_handle": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), c_lp, c_lp]),
    "scmc_multi_randomize": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "scmc_multi_set_stream_version": (C.c_int32, [C.c_void_p, C.c_int32]),
    "scmc_multi_set_sweep_mode": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64]),
    "scmc_multi_sweep": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, c_dp, c_dp, C.c_uint64, C.c_uint64, c_lp, c_lp]),
    "scmc_multi_energy": (C.c_int32, [C.c_void_p, c_dp]),
    "scmc_multi_optimize": (C.c_int32, [C.c_void_p, C.c_int64, c_dp]),
    "scmc_multi_get_state": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, c_dp, c_dp]),
    "scmc_multi_measure": (C.c_int32, [C.c_void_p, c_dp]),
    "scmc_multi_run_metropolis": (C.c_int32, [C.c_void_p, C.POINTER(RunParams), C.POINTER(RunResults)]),
    "scmc_multi_anneal": (C.c_int32, [C.c_void_p, C.POINTER(AnnealParams), C.POINTER(AnnealResults)]),
}

_lib = None


def load():
    """Load the CUDA library; raises ScmcError when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScmcError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise ScmcError(f"libscmc_b200 status {status}: {load().scmc_last_error().decode(errors='replace')}")


def dptr(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def lptr(a):
    return a.ctypes.data_as(c_lp) if a is not None else None


def bptr(a):
    return a.ctypes.data_as(c_bp) if a is not None else None


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def device_count():
    n = C.c_int32(0)
    st = load().scmc_device_count(C.byref(n))
    return n.value if st == 0 else 0
