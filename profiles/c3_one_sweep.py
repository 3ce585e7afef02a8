"""A few sweeps of BASELINE config C3 (honeycomb 96x96, n = 2, 256 replicas, FP32) for ncu: python profiles/c3_one_sweep.py [onsite]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
R = 256
Js = [np.eye(16)] if len(sys.argv) < 2 else [np.diag(np.r_[0.0, np.linspace(-0.2, 0.2, 15)]), np.eye(16)]
cfg = scmc.initializeCfg("nearest-neighbor", "honeycomb", Js, 96, 2, n_replicas=R, precision=32, seed=1)
cfg.set_lanes(1)
cfg.sweep(3, np.full(R, 5.0), np.full(R, 0.2), want_counts=False)
print("ok", cfg.info())
