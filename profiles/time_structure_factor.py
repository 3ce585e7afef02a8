"""Structure factor on allowed momenta: direct sum per momentum against the separable DFT over the cell indices, and against a sweep.
python profiles/time_structure_factor.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
def timed(f, reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)), out
for lat, L, n, R, box, nk_direct in (("honeycomb", 24, 2, 64, 4 * np.pi, None), ("honeycomb", 96, 2, 256, 4 * np.pi, 2048), ("triangular", 48, 1, 64, 4 * np.pi, None), ("triangular", 128, 1, 64, 4 * np.pi, 2048)):
    J = [np.eye(16)]
    cfg = scmc.initializeCfg("nearest-neighbor", lat, J, L, n, n_replicas=R, precision=32, seed=1)
    ks = scmc.getKsInBox(lat, L, [box, box], [0.0, 0.0])
    beta = np.full(R, 5.0); sigma = np.full(R, 0.2)
    cfg.sweep(10, beta, sigma, want_counts=False)
    t_sw, _ = timed(lambda: cfg.sweep(20, beta, sigma, want_counts=False), 3); t_sw /= 20
    sub = ks if nk_direct is None else ks[:nk_direct]
    scmc.structureFactor(cfg, sub[:64], method="direct"); scmc.structureFactor(cfg, sub[:64], method="grid")
    t_d, Sd = timed(lambda: scmc.structureFactor(cfg, sub, method="direct"), 3)
    t_g, Sg = timed(lambda: scmc.structureFactor(cfg, sub, method="grid"), 3)
    t_gall, Sall = timed(lambda: scmc.structureFactor(cfg, ks, method="grid"), 3)
    t_sum, Ssum = timed(lambda: scmc.structureFactorSum(cfg, ks), 3)
    plan = scmc.StructureFactorPlan(cfg, ks); plan.accumulate()
    t_acc, _ = timed(lambda: plan.accumulate(), 5); plan.mean()
    err = np.abs(Sg - Sd).max() / np.abs(Sd).max()
    print(f"{lat} L={L} n={n} R={R}: N={cfg.nSites}, {len(ks)} allowed k in the box; sweep {1e3*t_sw:.3f} ms | {len(sub)} k: direct {1e3*t_d:8.1f} ms, grid {1e3*t_g:7.1f} ms ({t_d/t_g:6.1f}x), "
          f"max rel. diff {err:.1e} | all {len(ks)} k through the grid path: {1e3*t_gall:8.1f} ms = {t_gall/t_sw:7.1f} sweeps (output {Sall.nbytes/2**20:.0f} MiB; direct would take ~{1e3*t_d*len(ks)/len(sub):.0f} ms); summed over components 0..14 on the device: {1e3*t_sum:7.1f} ms = {t_sum/t_sw:6.1f} sweeps (output {Ssum.nbytes/2**20:.0f} MiB); accumulated on the device (no copy): {1e3*t_acc:7.2f} ms = {t_acc/t_sw:6.1f} sweeps", flush=True)
    cfg.close()
