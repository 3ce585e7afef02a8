"""Phases of run! for one TG/h-BN shape: python profiles/time_tghbn_run_phases2.py L R"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
L, R = int(sys.argv[1]), int(sys.argv[2])
cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], L, 2, n_replicas=R, precision=32, seed=3)
Ts = np.geomspace(0.02, 2.0, R)
def timed(label, f):
    torch.cuda.synchronize(); t = time.perf_counter(); l0 = cfg.info()["launches"]
    f()
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"  {label:66s} {1e3*dt:8.2f} ms, {cfg.info()['launches'] - l0:5d} launches", flush=True)
beta = 1.0 / Ts; sigma = np.full(R, 0.3)
print(f"tg-hbn L={L} R={R} cluster {cfg.info()['cluster']}")
timed("warm-up: 20 sweeps", lambda: cfg.sweep(20, beta, sigma, want_counts=False))
timed("500 sweeps in one call (auto mode)", lambda: cfg.sweep(500, beta, sigma, want_counts=False))
timed("50 measurements", lambda: [cfg.measure() for _ in range(50)])
timed("run! therm 500, rate 10 (50 log + adaptation points)", lambda: scmc.run_(cfg, None, Ts, 2.0, 500, 0, measurement_rate=10))
timed("run! therm 500, rate 1000 (no points)", lambda: scmc.run_(cfg, None, Ts, 2.0, 500, 0, measurement_rate=1000))
timed("run! measurement phase only, 500 sweeps, rate 10 (50 rows)", lambda: scmc.run_(cfg, None, Ts, 2.0, 500, 500, measurement_rate=10, current_sweep=500))
timed("run! measurement phase only, 500 sweeps, rate 1000 (no rows)", lambda: scmc.run_(cfg, None, Ts, 2.0, 500, 500, measurement_rate=1000, current_sweep=500))
