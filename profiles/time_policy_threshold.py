"""Is "all clusters on the chip at once" too strict?  Shapes just above it, auto mode (host-driven streaming) against the forced resident
mode (device-side schedule), run! 500 + 500, 5 repeats.  python profiles/time_policy_threshold.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
tg = lambda L, n, R: (lambda: scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], L, n, n_replicas=R, precision=32, seed=3))
nn = lambda lat, L, n, R, Js: (lambda: scmc.initializeCfg("nearest-neighbor", lat, Js, L, n, n_replicas=R, precision=32, seed=3))
cases = [("tg-hbn L=48 n=2 x64", tg(48, 2, 64), 2), ("tg-hbn L=48 n=2 x128", tg(48, 2, 128), 2), ("tg-hbn L=48 n=2 x256", tg(48, 2, 256), 2), ("tg-hbn L=48 n=1 x128", tg(48, 1, 128), 2),
         ("tri 128 n=1 x64 (psi)", nn("triangular", 128, 1, 64, [J]), 4), ("tri 128 n=1 x148 (psi)", nn("triangular", 128, 1, 148, [J]), 4), ("tri 128 n=1 x512 (psi)", nn("triangular", 128, 1, 512, [J]), 4)]
for name, mk, forced in cases:
    for mode in (0, forced):
        cfg = mk(); R = cfg.n_replicas; cfg.set_sweep_mode(mode)
        Ts = np.geomspace(0.02, 2.0, R)
        ts = []
        for rep in range(5):
            l0 = cfg.info()["launches"]
            torch.cuda.synchronize(); t = time.perf_counter()
            scmc.run_(cfg, None, Ts, 2.0, 500, 500, measurement_rate=10)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t)); nl = cfg.info()["launches"] - l0
        print(f"{name:28s} cluster {cfg.info()['cluster']} x {R:4d} = {cfg.info()['cluster']*R:5d} CTAs, mode {mode}: median {np.median(ts):7.1f} ms (min {min(ts):7.1f}), launches {nl}", flush=True)
        cfg.close()
