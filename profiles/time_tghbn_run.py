"""TG/h-BN: full run! (host-driven schedule on the T-gather kernels) against the pure sweep time of the same handle.
python profiles/time_tghbn_run.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
for L, n, R in ((12, 2, 64), (12, 2, 1024), (24, 2, 64), (48, 2, 64), (48, 2, 1024)):
    cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], L, n, n_replicas=R, precision=32, seed=3)
    info = cfg.info()
    Ts = np.geomspace(0.02, 2.0, R)
    l0 = info["launches"]
    torch.cuda.synchronize(); t = time.perf_counter()
    scmc.run_(cfg, None, Ts, 2.0, 500, 500, measurement_rate=10)
    torch.cuda.synchronize(); t_run = time.perf_counter() - t
    nl = cfg.info()["launches"] - l0
    beta = 1.0 / Ts; sigma = np.full(R, 0.3)
    cfg.sweep(50, beta, sigma, want_counts=False)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(4): cfg.sweep(50, beta, sigma, want_counts=False)
    torch.cuda.synchronize(); t_sw = (time.perf_counter() - t) / 200
    t = time.perf_counter()
    for _ in range(20): cfg.measure()
    torch.cuda.synchronize(); t_me = (time.perf_counter() - t) / 20
    print(f"tg-hbn L={L} n={n} R={R:5d} resident {info['resident']} cluster {info['cluster']}: run! 1000 sweeps {1e3*t_run:7.1f} ms = {1e3*t_run:.1f} us/sweep incl. 100 measurements, "
          f"{nl} launches | 50-sweep calls {1e6*t_sw:6.1f} us/sweep | measure {1e6*t_me:6.1f} us | run! / (1000 sweeps + 100 measures) = {t_run/(1000*t_sw+100*t_me):.2f}", flush=True)
    cfg.close()
