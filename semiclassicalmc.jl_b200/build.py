"""Build recipe for libscmc_b200.so (hand-written sm_100a CUDA + C ABI).  In-tree output so the library travels
with the repo snapshot to the GPU box; nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscmc_b200.so")
SOURCES = ["scmc_core.cu", "scmc_drivers.cu", "scmc_resident.cu", "scmc_tma.cu", "scmc_multi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-use_fast_math",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    from . import generators
    gen_h = os.path.join(CSRC, "gen_code.h")
    if force or _stale(gen_h, [os.path.join(HERE, "generators.py")]):
        generators.emit_cuda_tables(gen_h)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "scmc.h"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run(["nvcc", *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        with open(o + ".log", "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or force or _stale(LIB, objs):
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(HERE, ".."))
    import scmc_b200  # noqa: F401  (resolves the relative import above)
    print(scmc_b200.build_library(force="--force" in sys.argv, verbose=True))
