"""Few replicas of a small lattice: does a larger cluster than needed (more SMs per replica) pay?  python profiles/bench_cluster_size.py
Chains do not depend on the cluster size (checked here before timing)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
for L, R in ((48, 64), (48, 16), (48, 148), (36, 64), (96, 32)):
    ref = None
    for cmin in (1, 2, 4):
        os.environ["SCMC_CLUSTER_MIN"] = str(cmin)
        cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [J], L, 1, n_replicas=R, precision=32, seed=5)
        info = cfg.info()
        if not info["resident_psi"] or info["cluster"] < cmin:
            cfg.close(); continue
        cfg.set_sweep_mode(4)
        beta = np.linspace(0.5, 20, R); sigma = np.full(R, 0.3)
        cfg.sweep(50, beta, sigma, want_counts=False)
        st = cfg.get_state()
        if ref is None: ref = st
        same = np.array_equal(ref, st)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(6):
            cfg.sweep(50, beta, sigma, want_counts=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 300
        print(f"L={L} R={R:4d} cluster {info['cluster']} (CTAs {R * info['cluster']:4d} of 148 SMs): {1e6*dt:7.1f} us/sweep, {2*L*L*R/dt:.3e} updates/s, chains identical to cluster 1: {same}", flush=True)
        cfg.close()
os.environ.pop("SCMC_CLUSTER_MIN", None)
