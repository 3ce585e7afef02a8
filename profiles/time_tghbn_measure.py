"""Time of one measurement against one sweep on TG/h-BN (L = 48, n = 2, 1024 replicas, FP32): python profiles/time_tghbn_measure.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
for n in (2, 1):
    R = 1024
    cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 48, n, n_replicas=R, precision=32, seed=1)
    cfg.set_sweep_mode(1)
    beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
    cfg.sweep(5, beta, sigma, want_counts=False); cfg.measure()
    torch.cuda.synchronize(); t = time.perf_counter()
    cfg.sweep(50, beta, sigma, want_counts=False)
    torch.cuda.synchronize(); t_sweep = (time.perf_counter() - t) / 50
    t = time.perf_counter()
    for _ in range(20):
        cfg.measure()
    torch.cuda.synchronize(); t_meas = (time.perf_counter() - t) / 20
    print(f"tg-hbn L=48 n={n} R={R}: sweep {1e6*t_sweep:.1f} us, measure (incl. 172 KB copy to host) {1e6*t_meas:.1f} us = {t_meas/t_sweep:.2f} sweeps; "
          f"at measurement_rate 10: {100*t_meas/(10*t_sweep+t_meas):.1f} % of a run")
    cfg.close()
