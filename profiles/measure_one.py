"""One psi-family measurement on the C5 lattice (for ncu): python profiles/measure_one.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
R = 2048
cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [J], 128, 1, n_replicas=R, precision=32, seed=1)
cfg.sweep(2, np.full(R, 5.0), np.full(R, 0.3), want_counts=False)
obs = cfg.measure(); obs = cfg.measure()
print("ok", obs[0, 0])
