"""Runs profiles/all_kernels_small.py: every kernel and every sweep mode (auto, streaming / resident x T-gather / psi-gather) on
five small models, including the drivers, observables and the injected-random path.  (compute-sanitizer is closed on this GPU
pool; table validation happens in scmc_create, and results are checked against the oracle in the other GPU tests.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_every_kernel_and_mode_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "all_kernels_small.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sanitize workload done" in out.stdout and out.stdout.count("ok ") >= 18
