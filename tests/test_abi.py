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
    assert ctypes.sizeof(L.AnnealResults) == 8 * 8


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


def _c_prototypes():
    text = open(os.path.join(ROOT, "include", "scmc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int32_t|const char\*)\s+(scmc_[A-Za-z_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return protos, text


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_julia_shim_binds_declared_symbols_with_matching_arity():
    """The Julia shim cannot run here (no Julia in the image), so it is checked statically against include/scmc.h: every ccall names a
    declared symbol, passes as many argument types as the C prototype has parameters and as many values as types; the Julia mirror
    structs list the header's fields in order; every on-path name the reference exports (src/SemiClassicalMC.jl:19-87) is defined."""
    protos, header = _c_prototypes()
    src = open(os.path.join(ROOT, "julia", "SemiClassicalMCB200.jl")).read()
    calls = 0
    for m in re.finditer(r"ccall\(\(:(scmc_[a-z_]+), LIB\),\s*(\w+),\s*\(", src):
        name = m.group(1)
        assert name in protos, f"{name} is not declared in include/scmc.h"
        # argument-type tuple: balanced parentheses starting at the end of the match
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0); i += 1
        types = [t for t in _split_top(src[m.end():i - 1]) if t.strip()]
        assert len(types) == protos[name], (name, len(types), protos[name])
        # values: up to the parenthesis closing the ccall
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0); j += 1
        values = [v for v in _split_top(src[i:j - 1].lstrip(", \n")) if v.strip()]
        assert len(values) == len(types), (name, len(values), len(types))
        calls += 1
    assert calls >= 25
    for struct, cname in (("RunParams", "scmc_run_params"), ("RunResults", "scmc_run_results"), ("AnnealParams", "scmc_anneal_params"),
                          ("AnnealResults", "scmc_anneal_results")):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + ";", header).group(1)
        c_fields = re.findall(r"(\w+)\s*(?:,|;)", re.sub(r"\b(const|double|int64_t|int32_t|uint64_t)\b|\*", " ", body))
        jl = re.search(r"struct " + struct + r"\n(.*?)\nend", src, flags=re.S).group(1)
        jl_fields = re.findall(r"(\w+)::", jl)
        assert [f.lower() for f in jl_fields] == [f.lower() for f in c_fields], (struct, jl_fields, c_fields)
    for name in ("Configuration", "initializeCfg", "getEnergy", "getSpinExpectation", "getState", "getMagnetization", "getCorrelations", "getRs",
                 "getProject", "initializeObservables", "measure!", "saveMeans!", "checkpoint!", "readCheckpoint", "readCheckpointSA", "run!",
                 "runAnnealing!", "localUpdate!", "localOptimization!", "launch!", "launchAnnealing!"):
        assert re.search(r"(function\s+" + re.escape(name) + r"\(|^" + re.escape(name) + r"\(.*\).*=|struct " + re.escape(name) + r"\b)", src, flags=re.M), name
