"""The 1318 ms row of time_tghbn_run.py, repeated on fresh handles in one process.  python profiles/time_tghbn_run_repeat.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
R = 1024
Ts = np.geomspace(0.02, 2.0, R)
for rep, (nth, nme) in enumerate([(500, 500), (500, 500), (200, 200), (500, 500)]):
    cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 48, 2, n_replicas=R, precision=32, seed=3)
    torch.cuda.synchronize(); t = time.perf_counter(); l0 = cfg.info()["launches"]
    res = scmc.run_(cfg, None, Ts, 2.0, nth, nme, measurement_rate=10)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"rep {rep}: fresh handle, run! {nth}+{nme}: {1e3*dt:8.1f} ms = {1e6*dt/(nth+nme):7.1f} us/sweep, {cfg.info()['launches'] - l0} launches, "
          f"acc rate min/max {res['acc_rate'].min():.3f}/{res['acc_rate'].max():.3f}, sigma min/max {res['sigma'].min():.3f}/{res['sigma'].max():.2f}", flush=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    cfg.sweep(50, 1.0 / Ts, res["sigma"], want_counts=False)
    torch.cuda.synchronize(); print(f"        50 sweeps afterwards with the adapted sigma: {1e6*(time.perf_counter() - t)/50:7.1f} us/sweep", flush=True)
    cfg.close()
