"""Full run! (ramp, sigma adaptation, measurements every 10 sweeps) on small workloads with the cluster size forced:
cluster 1 = device-side schedule inside the persistent kernel, larger clusters = host-driven schedule with resident sweeps.
python profiles/bench_cluster_run.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc
import torch
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
cases = [("C2 tri 48 x64T", "triangular", 48, 64, [J], 1000, 1000), ("tri 48 x16T", "triangular", 48, 16, [J], 1000, 1000),
         ("C1 honeycomb 6 x1", "honeycomb", 6, 1, [np.eye(16)], 5000, 5000), ("tri 36 x64", "triangular", 36, 64, [J], 1000, 1000)]
for name, lat, L, R, Js, nth, nme in cases:
    ref = None
    for cmin in (1, 2, 4, 8):
        os.environ["SCMC_CLUSTER_MIN"] = str(cmin)
        cfg = scmc.initializeCfg("nearest-neighbor", lat, Js, L, 1, n_replicas=R, precision=32, seed=5)
        info = cfg.info()
        if not info["resident_psi"] or info["cluster"] < cmin:
            cfg.close(); continue
        Ts = np.geomspace(0.02, 2.0, R) if R > 1 else np.array([0.1])
        l0 = info["launches"]
        torch.cuda.synchronize(); t = time.perf_counter()
        res = scmc.run_(cfg, None, Ts, 2.0, nth, nme, measurement_rate=10)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        st = cfg.get_state()
        if ref is None: ref = (st, res["mean"][:, 0].copy())
        print(f"{name:20s} cluster {info['cluster']}: run! {nth}+{nme} sweeps {1e3*dt:8.1f} ms = {1e6*dt/(nth+nme):6.1f} us/sweep, launches {cfg.info()['launches'] - l0:5d}, "
              f"states identical: {np.array_equal(ref[0], st)}, max |d<E/N>| {np.abs(res['mean'][:, 0] - ref[1]).max():.1e}", flush=True)
        cfg.close()
os.environ.pop("SCMC_CLUSTER_MIN", None)
