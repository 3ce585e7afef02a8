"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/scmc.h declares, the ctypes
binding covers them all, and compute entries fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "scmc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scmc_[A-Za-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(scmc):
    lib = scmc._lib.load()
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/scmc.h but not exported"
    assert set(syms) == set(scmc._lib.SIGNATURES), set(syms) ^ set(scmc._lib.SIGNATURES)
    assert lib.scmc_version() >= 100


def test_struct_layouts_match_header(scmc):
    L = scmc._lib
    assert ctypes.sizeof(L.RunParams) == 8 * 3 + 4 * 2 + 8 * 2 + 8 * 3
    assert ctypes.sizeof(L.RunResults) == 8 * 5
    assert ctypes.sizeof(L.AnnealParams) == 8 * 8 + 4 * 2 + 8
    assert ctypes.sizeof(L.AnnealResults) == 8 * 4


def test_no_cpu_fallback(scmc):
    """Without a GPU, creating a Configuration must raise (the product never computes on the CPU)."""
    if scmc._lib.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(scmc._lib.ScmcError, match="CUDA"):
        scmc.initializeCfg("nearest-neighbor", "triangular", [np.eye(16)], 3, 1)


def test_argument_validation_before_any_cuda_call(scmc):
    lib = scmc._lib.load()
    h = ctypes.c_void_p()
    st = lib.scmc_create(ctypes.byref(h), 0, 16, 9, 4, 1, 0, 6, None, None, 1, None, None, None, None, None, None, None)
    assert st == -1 and b"precision" in lib.scmc_last_error()
    assert lib.scmc_destroy(None) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or the C sources may reference it."""
    pkg = os.path.join(ROOT, "semiclassicalmc.jl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, os.path.join(base, f)


def _build_client(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_client")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "abi_client.c"), "-ldl"])
    return exe


def test_plain_c_client_binds_the_abi(scmc, tmp_path):
    """A plain-C program (what a cgo / ccall binding amounts to) loads the library and gets clean error codes without a GPU."""
    import subprocess
    exe = _build_client(tmp_path)
    out = subprocess.run([exe, scmc._lib.LIB_PATH], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad precision -> status -1" in out.stdout


@pytest.mark.gpu
def test_plain_c_client_runs_a_simulation(scmc, tmp_path):
    import subprocess
    exe = _build_client(tmp_path)
    out = subprocess.run([exe, scmc._lib.LIB_PATH, "run"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "abi_client ok" in out.stdout, out.stdout + out.stderr
