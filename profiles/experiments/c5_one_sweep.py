"""A few sweeps of the C5 lattice (triangular 128x128, n = 1, 2048 replicas, FP32, streaming psi-gather kernel) for ncu:
    SCMC_MMA=1 python profiles/experiments/c5_one_sweep.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import scmc_b200 as scmc
R = 2048
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [J], 128, 1, n_replicas=R, precision=32, seed=1, stream=int(os.environ.get("STREAM", "2")))
cfg.set_sweep_mode(int(os.environ.get("SCMC_MODE", "3"))); cfg.set_lanes(1)
T = np.tile(np.geomspace(0.02, 2.0, 64), R // 64)
cfg.sweep(2, 1 / T, np.clip(0.6 * np.sqrt(T), 0.02, 60), want_counts=False)
print("ok", cfg.info())
