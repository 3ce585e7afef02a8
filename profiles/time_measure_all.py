"""Time of one measurement (scmc_measure, host buffers) against one sweep for the T-gather-family workloads and the headline one:
python profiles/time_measure_all.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
cases = [("tg-hbn L=48 n=2 x1024", lambda: scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 48, 2, n_replicas=1024, precision=32, seed=1), 1),
         ("tg-hbn L=48 n=1 x1024", lambda: scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 48, 1, n_replicas=1024, precision=32, seed=1), 1),
         ("C3 honeycomb 96 n=2 x256 (T family)", lambda: scmc.initializeCfg("nearest-neighbor", "honeycomb", [np.eye(16)], 96, 2, n_replicas=256, precision=32, seed=1), 0),
         ("C2 tri 48 n=1 x64 (psi family)", lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [J], 48, 1, n_replicas=64, precision=32, seed=1), 3),
         ("tri 128 n=1 x1024 (psi family)", lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [J], 128, 1, n_replicas=1024, precision=32, seed=1), 0),
         ("tri 36 n=1 x512 FP64 mode 1 (T family)", lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [J], 36, 1, n_replicas=512, precision=64, seed=1), 1)]
for name, mk, mode in cases:
    cfg = mk(); cfg.set_sweep_mode(mode)
    R = cfg.n_replicas
    beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
    cfg.sweep(5, beta, sigma, want_counts=False); cfg.measure()
    torch.cuda.synchronize(); t = time.perf_counter()
    cfg.sweep(40, beta, sigma, want_counts=False)
    torch.cuda.synchronize(); t_sweep = (time.perf_counter() - t) / 40
    t = time.perf_counter()
    for _ in range(20):
        obs = cfg.measure()
    torch.cuda.synchronize(); t_meas = (time.perf_counter() - t) / 20
    print(f"{name:40s} sweep {1e6*t_sweep:8.1f} us  measure {1e6*t_meas:8.1f} us = {t_meas/t_sweep:5.2f} sweeps   E/N[0] = {obs[0,0]:+.9f}", flush=True)
    cfg.close()
