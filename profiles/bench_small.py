"""Times colour-ordered sweeps for the smaller BASELINE workloads (C1, C2, C4 shapes) with the streaming (modes 1, 3) and the
replica-resident (modes 2, 4) kernels of both arithmetic families: python profiles/bench_small.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch

J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
cases = [("C1 honeycomb 6x6", "honeycomb", 6, 1, 4096, [np.eye(16)]), ("C1 single", "honeycomb", 6, 1, 1, [np.eye(16)]),
         ("C2 tri 48x48 x64T", "triangular", 48, 1, 64, [J]), ("C4 tri 36x36 x512", "triangular", 36, 1, 512, [J]),
         ("C2 tri 48x48 x4096", "triangular", 48, 1, 4096, [J]),
         ("C3 honeycomb 96x96 n=2 x256", "honeycomb", 96, 2, 256, [np.eye(16)]),
         ("C3 n=2 with on-site K", "honeycomb", 96, 2, 256, [np.diag(np.r_[0.0, np.linspace(-0.2, 0.2, 15)]), np.eye(16)])]
for name, lat, L, n, R, Js in cases:
    for mode in (1, 2, 3, 4):
        cfg = scmc.initializeCfg("nearest-neighbor", lat, Js, L, n, n_replicas=R, precision=32, seed=1)
        info = cfg.info()
        if (mode == 2 and not info["resident"]) or (mode == 4 and not info["resident_psi"]):
            cfg.close()
            continue
        cfg.set_sweep_mode(mode)
        beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
        for sweeps_per_call in (1, 50):
            cfg.sweep(sweeps_per_call, beta, sigma, want_counts=False)
            torch.cuda.synchronize(); t = time.perf_counter()
            calls = max(1, 200 // sweeps_per_call)
            for _ in range(calls):
                cfg.sweep(sweeps_per_call, beta, sigma, want_counts=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            ups = calls * sweeps_per_call * 2 * cfg.nSites * R / dt
            print(f"{name:22s} mode {mode} cluster {info['cluster']} sweeps/call {sweeps_per_call:3d}: {ups:.3e} updates/s, {1e6*dt/(calls*sweeps_per_call):8.1f} us/sweep", flush=True)
        cfg.close()
