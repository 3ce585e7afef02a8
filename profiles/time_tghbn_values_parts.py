"""Parts of ObservablesTgHbn.values (median of 20): python profiles/time_tghbn_values_parts.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import scmc_b200.observables as ob
import torch
def med(f, n=20):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return 1e6 * float(np.median(ts))
for L, R in ((24, 64), (24, 1024)):
    cfg = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], L, 2, n_replicas=R, precision=32, seed=3)
    obs = scmc.initializeObservables(scmc.ObservablesTgHbn, cfg)
    cfg.sweep(5, np.full(R, 5.0), np.full(R, 0.2), want_counts=False); obs.values(cfg)
    tri = np.ascontiguousarray(cfg.interactionSites[:, [0, 4, 3, 1]], dtype=np.int32)
    subM = ob.getSublatticeMagnetization(cfg, obs.siteLabels)
    def post():
        m = subM.mean(axis=2); out = {}
        for key, comps in ob._TGHBN_SETS.items():
            out["subM" + key] = np.linalg.norm(subM[:, comps, :], axis=1).mean(axis=1); out["m" + key] = np.linalg.norm(m[:, comps], axis=1)
    print(f"tg-hbn L={L} R={R}: values {med(lambda: obs.values(cfg)):7.1f} us = chirality {med(lambda: ob.getChirality(cfg)):6.1f} (with a prepared table {med(lambda: ob.getChirality(cfg, tri)):6.1f}) "
          f"+ sublattice magnetisation {med(lambda: ob.getSublatticeMagnetization(cfg, obs.siteLabels)):6.1f} + numpy post-processing {med(post):6.1f}; cfg.measure {med(lambda: cfg.measure()):6.1f}; "
          f"empty API call (scmc_info) {med(lambda: cfg.info()):5.1f}", flush=True)
    cfg.close()
