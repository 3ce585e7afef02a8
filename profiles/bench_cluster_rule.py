"""Chip-filling cluster rule (scmc_create) on and off (SCMC_CLUSTER_FILL=0): full run!, annealing and plain sweeps in auto mode for few replicas.
python profiles/bench_cluster_rule.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
cases = [("C2 tri 48 x64T", "triangular", 48, 64, [J]), ("tri 48 x16", "triangular", 48, 16, [J]), ("C1 honeycomb 6 x1", "honeycomb", 6, 1, [np.eye(16)]),
         ("tri 36 x64", "triangular", 36, 64, [J]), ("tri 36 x512 (C4 shard)", "triangular", 36, 512, [J]), ("tri 96 x8", "triangular", 96, 8, [J]), ("tri 128 x4", "triangular", 128, 4, [J])]
for name, lat, L, R, Js in cases:
    ref = None
    for fill in ("0", "1"):
        os.environ["SCMC_CLUSTER_FILL"] = fill
        mk = lambda: scmc.initializeCfg("nearest-neighbor", lat, Js, L, 1, n_replicas=R, precision=32, seed=5)
        cfg = mk(); cl = cfg.info()["cluster"]
        Ts = np.geomspace(0.02, 2.0, R) if R > 1 else np.array([0.1])
        torch.cuda.synchronize(); t = time.perf_counter()
        res = scmc.run_(cfg, None, Ts, 2.0, 500, 500, measurement_rate=10)
        torch.cuda.synchronize(); t_run = time.perf_counter() - t
        st = cfg.get_state()
        # plain sweeps through the public call, auto mode
        beta = 1.0 / Ts; sigma = np.full(R, 0.3)
        cfg.sweep(20, beta, sigma, want_counts=False)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): cfg.sweep(20, beta, sigma, want_counts=False)
        torch.cuda.synchronize(); t_sw = (time.perf_counter() - t) / 100
        cfg.close()
        cfg = mk()
        torch.cuda.synchronize(); t = time.perf_counter()
        an = scmc.runAnnealing_(cfg, 2.0, T_fac=0.9, N_max=600, N_o=0, N_per_T=20, min_acc_per_site=3, min_accrate=0.02, σ_min=0.05)
        torch.cuda.synchronize(); t_an = time.perf_counter() - t
        st_an = cfg.get_state(); cfg.close()
        if ref is None: ref = (st, st_an)
        print(f"{name:24s} fill {fill} cluster {cl}: run! 500+500 {1e3*t_run:7.1f} ms | sweep (auto) {1e6*t_sw:7.1f} us | anneal ({int(an['sweeps'].max())} sweeps) {1e3*t_an:7.1f} ms | "
              f"identical to fill 0: {np.array_equal(ref[0], st) and np.array_equal(ref[1], st_an)}", flush=True)
os.environ.pop("SCMC_CLUSTER_FILL", None)
