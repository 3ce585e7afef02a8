"""Optimisation sweep (localOptimization!, the gradient-descent half of C4) against a Metropolis sweep at the C4 shard shape (triangular 36x36, 512
restarts) and two others; median of 5.  python profiles/time_optimize.py"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
from scmc_b200 import _lib
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
def med(f, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return float(np.median(ts))
for name, mk in (("C4 shard: tri 36 n=1 x512", lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [J], 36, 1, n_replicas=512, precision=32, seed=1)),
                 ("tri 36 n=1 x512 FP64", lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [J], 36, 1, n_replicas=512, precision=64, seed=1)),
                 ("tg-hbn 24 n=2 x64", lambda: scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 24, 2, n_replicas=64, precision=32, seed=1))):
    cfg = mk(); R = cfg.n_replicas
    beta = np.full(R, 20.0); sigma = np.full(R, 0.1)
    cfg.sweep(50, beta, sigma, want_counts=False)
    t_sw = med(lambda: cfg.sweep(50, beta, sigma, want_counts=False)) / 50
    E = np.zeros(R)
    opt = lambda n: _lib.check(cfg._lib.scmc_optimize(cfg._h, n, _lib.dptr(E)))
    opt(2)
    t_o10 = med(lambda: opt(10)); t_o1 = med(lambda: opt(1))
    print(f"{name:28s}: Metropolis sweep {1e6*t_sw:7.1f} us | optimisation: 1 sweep per call {1e6*t_o1:8.1f} us, 10 per call {1e6*t_o10/10:8.1f} us per sweep = {t_o10/10/t_sw:5.1f} Metropolis sweeps; E/N after = {E.mean()/cfg.nSites:+.6f}", flush=True)
    cfg.close()
