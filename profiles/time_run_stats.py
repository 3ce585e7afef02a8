"""Full run! (500 + 500 sweeps, a measurement every 10) repeated 5x per workload on one handle: min / median / max, so that single host stalls
(cudaMalloc / cudaFree tails, profiles/README.md) do not decide a comparison.  python profiles/time_run_stats.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
tg = lambda L, n, R: (lambda: scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], L, n, n_replicas=R, precision=32, seed=3))
nn = lambda lat, L, n, R, Js: (lambda: scmc.initializeCfg("nearest-neighbor", lat, Js, L, n, n_replicas=R, precision=32, seed=3))
cases = [("tg-hbn L=12 n=2 x64", tg(12, 2, 64)), ("tg-hbn L=12 n=2 x1024", tg(12, 2, 1024)), ("tg-hbn L=24 n=2 x64", tg(24, 2, 64)), ("tg-hbn L=48 n=2 x64", tg(48, 2, 64)),
         ("tg-hbn L=48 n=2 x1024", tg(48, 2, 1024)), ("tg-hbn L=24 n=1 x16", tg(24, 1, 16)), ("honeycomb 24 n=2 x64 (T family)", nn("honeycomb", 24, 2, 64, [np.eye(16)])),
         ("C2 tri 48 n=1 x64 (psi family)", nn("triangular", 48, 1, 64, [J])), ("C1 honeycomb 6 n=1 x1 (psi family)", nn("honeycomb", 6, 1, 1, [np.eye(16)]))]
for name, mk in cases:
    cfg = mk(); R = cfg.n_replicas
    Ts = np.geomspace(0.02, 2.0, R) if R > 1 else np.array([0.1])
    ts, nl = [], 0
    for rep in range(5):
        l0 = cfg.info()["launches"]
        torch.cuda.synchronize(); t = time.perf_counter()
        scmc.run_(cfg, None, Ts, 2.0, 500, 500, measurement_rate=10)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t)); nl = cfg.info()["launches"] - l0
    ts = np.array(ts)
    print(f"{name:36s} cluster {cfg.info()['cluster']}: run! 500+500 ms min {ts.min():7.1f} median {np.median(ts):7.1f} max {ts.max():7.1f}   launches {nl}", flush=True)
    cfg.close()
