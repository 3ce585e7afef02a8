import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import scmc_b200 as scmc
R = 1024
cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 48, 2, n_replicas=R, precision=32, seed=1)
cfg.set_sweep_mode(1); cfg.set_lanes(1)
beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
cfg.sweep(3, beta, sigma, want_counts=False)
print("ok", cfg.info())
