import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import scmc_b200 as scmc
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
rng = np.random.default_rng(3)
bad = 0
for lattice, L in [("triangular", 40), ("triangular", 41), ("triangular", 64), ("triangular", 100), ("triangular", 130), ("honeycomb", 33), ("honeycomb", 64), ("square", 46), ("square", 64), ("triangular", 255)]:
    for R in (1, 33, 130):
        if L > 200 and R > 33: continue
        A = rng.uniform(-1, 1, (16, 16)); Jd = 0.5 * (A + A.T); Jd[0, 0] = 0
        out = {}
        for mode in (3, 5, 6):
            cfg = scmc.initializeCfg("nearest-neighbor", lattice, [Jd], L, 1, n_replicas=R, precision=32, seed=11, stream=2)
            cfg.set_sweep_mode(mode)
            info = cfg.info()
            nv = cfg.verify_tables() if mode == 6 else 0
            beta = np.geomspace(0.5, 30.0, R); sigma = np.geomspace(0.7, 0.05, R)
            acc, att = cfg.sweep(2, beta, sigma, hits=2)
            out[mode] = (acc.copy(), cfg.get_state(), cfg.energies().copy(), info["tma"], nv)
            cfg.close()
        same35 = np.array_equal(out[3][0], out[5][0]) and np.array_equal(out[3][1], out[5][1])
        dacc = np.abs(out[5][0].astype(np.int64) - out[6][0].astype(np.int64)).max()
        frac = (np.abs(out[5][1] - out[6][1]).max(axis=2) > 1e-4).mean()
        ok = same35 and dacc <= 6 and frac < 5e-4 and out[6][4] == 0
        bad += not ok
        print(f"{lattice} L={L} R={R} tma={out[6][3]} mode5==mode3 {same35}  |dacc| {dacc}  differing sites {frac:.1e}  verify {out[6][4]}  {'ok' if ok else 'BAD'}", flush=True)
print("fuzz", "ok" if bad == 0 else f"{bad} BAD")
