"""One device-side run! of BASELINE config C2 (triangular 48x48, n = 1, 64 temperatures; clusters of 2) for ncu: python profiles/c2_run_chunk.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
R = 64
cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [J], 48, 1, n_replicas=R, precision=32, seed=1)
res = scmc.run_(cfg, None, np.geomspace(0.02, 2.0, R), 2.0, 100, 200, measurement_rate=10)
print("ok", cfg.info(), res["rows"])
