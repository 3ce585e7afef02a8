"""TG/h-BN: cost of one measureTgHbn_ (generic row + sublattice magnetisations + chiralities + host-side binning per replica) against a sweep.
python profiles/time_tghbn_observables.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
for L, R in ((12, 64), (24, 64), (24, 1024), (48, 64)):
    cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], L, 2, n_replicas=R, precision=32, seed=3)
    obs = scmc.initializeObservables(scmc.ObservablesTgHbn, cfg)
    beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
    cfg.sweep(20, beta, sigma, want_counts=False); scmc.measureTgHbn_(obs, cfg)
    torch.cuda.synchronize(); t = time.perf_counter()
    cfg.sweep(50, beta, sigma, want_counts=False)
    torch.cuda.synchronize(); t_sw = (time.perf_counter() - t) / 50
    def med(f, n=10):
        ts = []
        for _ in range(n):
            torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        return float(np.median(ts))
    t_all = med(lambda: scmc.measureTgHbn_(obs, cfg))
    t_row = med(lambda: cfg.measure())
    t_vals = med(lambda: obs.values(cfg))
    t_sub = med(lambda: scmc.getSublatticeMagnetization(cfg, obs.sublattice_labels) if hasattr(obs, "sublattice_labels") else None)
    print(f"tg-hbn L={L} R={R:5d}: sweep {1e6*t_sw:7.1f} us | measureTgHbn_ {1e6*t_all:9.1f} us = {t_all/t_sw:6.1f} sweeps  [generic row {1e6*t_row:7.1f} us, device reductions obs.values {1e6*t_vals:8.1f} us, "
          f"host-side binning etc. {1e6*(t_all - t_row - t_vals):9.1f} us]", flush=True)
    cfg.close()
