"""Sweep throughput of the TG/h-BN model (3 coupling labels, 12 neighbour slots, on-site Hund term at n = 2):
python profiles/bench_tghbn.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch

Jtg = [1.0, 0.3, 0.2, 0.1, 0.5]
for L, n, R in ((48, 2, 1024), (48, 1, 1024), (12, 2, 4096), (96, 2, 256)):
    for mode in (0, 1, 2, 3):
        cfg = scmc.initializeCfg("tg-hbn", "triangular", Jtg, L, n, n_replicas=R, precision=32, seed=1)
        info = cfg.info()
        if (mode == 2 and not info["resident"]) or (mode == 3 and n == 2):
            cfg.close()
            continue
        cfg.set_sweep_mode(mode)
        beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
        for sweeps_per_call in (1, 20):
            cfg.sweep(sweeps_per_call, beta, sigma, want_counts=False)
            torch.cuda.synchronize(); t = time.perf_counter()
            calls = max(1, 60 // sweeps_per_call)
            for _ in range(calls):
                cfg.sweep(sweeps_per_call, beta, sigma, want_counts=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            ups = calls * sweeps_per_call * 2 * cfg.nSites * R / dt
            print(f"tg-hbn L={L} n={n} R={R} colours {info['n_colours']} mode {mode} cluster {info['cluster']} sweeps/call {sweeps_per_call:3d}: "
                  f"{ups:.3e} updates/s, {1e6*dt/(calls*sweeps_per_call):8.1f} us/sweep", flush=True)
        cfg.close()
