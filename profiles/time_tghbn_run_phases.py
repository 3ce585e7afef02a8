"""Where does a host-driven run! spend its time?  TG/h-BN L = 48, n = 2, 1024 replicas.  python profiles/time_tghbn_run_phases.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 48, 2, n_replicas=R, precision=32, seed=3)
Ts = np.geomspace(0.02, 2.0, R)
def timed(label, f):
    torch.cuda.synchronize(); t = time.perf_counter(); l0 = cfg.info()["launches"]
    f()
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{label:70s} {1e3*dt:8.1f} ms, {cfg.info()['launches'] - l0:5d} launches", flush=True)
beta = 1.0 / Ts; sigma = np.full(R, 0.3)
timed("warm-up: 20 sweeps in one call", lambda: cfg.sweep(20, beta, sigma, want_counts=False))
timed("200 sweeps in one call", lambda: cfg.sweep(200, beta, sigma, want_counts=False))
timed("200 sweeps, one call each", lambda: [cfg.sweep(1, beta, sigma, want_counts=False) for _ in range(200)])
timed("200 sweeps, 20 calls of 10", lambda: [cfg.sweep(10, beta, sigma, want_counts=False) for _ in range(20)])
timed("20 measurements", lambda: [cfg.measure() for _ in range(20)])
timed("run! therm 200 (ramp 150 one by one), rate 10, no measurement phase", lambda: scmc.run_(cfg, None, Ts, 2.0, 200, 0, measurement_rate=10))
timed("run! therm 200, rate 1000 (no adaptation / log points)", lambda: scmc.run_(cfg, None, Ts, 2.0, 200, 0, measurement_rate=1000))
timed("run! measurement phase only: 200 sweeps, rate 10", lambda: scmc.run_(cfg, None, Ts, 2.0, 200, 200, measurement_rate=10, current_sweep=200))
timed("run! measurement phase only: 200 sweeps, rate 10, keep_series=False", lambda: scmc.run_(cfg, None, Ts, 2.0, 200, 200, measurement_rate=10, current_sweep=200, keep_series=False))
