"""Small workload that touches every kernel of libscmc_b200 (every sweep mode on five models; compute-sanitizer is closed on this GPU pool, so this runs plain as a GPU test)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import scmc_b200 as scmc

J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
rng = np.random.default_rng(0)
A = rng.uniform(-1, 1, (16, 16)); Jd = 0.5 * (A + A.T)
cases = [("nearest-neighbor", "triangular", [J], 6, 1, 32), ("nearest-neighbor", "triangular", [J], 48, 1, 32), ("nearest-neighbor", "honeycomb", [np.diag(np.linspace(-.2, .2, 16)), Jd], 4, 2, 64),
         ("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.15, 0.8], 6, 2, 32), ("nearest-neighbor", "square", [Jd], 5, 1, 32)]
for model, lat, Js, L, n, prec in cases:
    for mode in (0, 1, 2, 3, 4):
        cfg = scmc.initializeCfg(model, lat, Js, L, n, n_replicas=3, precision=prec, seed=5)
        info = cfg.info()
        if (mode == 2 and not info["resident"]) or (mode == 4 and not info["resident_psi"]) or (mode in (3, 4) and len(cfg.interactions) > 1):
            cfg.close(); continue
        cfg.set_sweep_mode(mode)
        cfg.sweep(2, [1.0, 2.0, 4.0], [0.5, 0.4, 0.3])
        cfg.set_state(cfg.get_state())
        cfg.energies(); cfg.measure(); cfg.local_field(1); cfg.get_T(2, squared=True)
        scmc.run_(cfg, None, [0.5, 0.3, 0.2], 1.0, 12, 12, measurement_rate=4)
        scmc.runAnnealing_(cfg, 2.0, T_fac=0.8, N_max=40, N_o=2, N_per_T=5, min_acc_per_site=2, min_accrate=0.01)
        if mode == 0:
            B = 2 * np.pi * np.linalg.inv(np.stack(cfg.unitVectors) * cfg.L).T
            ks = np.array([B[0], B[1], B[0] + B[1]])
            scmc.structureFactor(cfg, ks)                       # allowed momenta: separable DFT over the cell indices (2-D lattices)
            scmc.structureFactorSum(cfg, ks)                    # summed over components on the device
            scmc.structureFactor(cfg, ks + 1e-3)                # arbitrary momenta: direct sum per momentum
            if cfg.nSites <= 80:
                scmc.getCorrelations(cfg, scmc.getProject(cfg), 1)
            if model == "tg-hbn":
                obs = scmc.ObservablesTgHbn(cfg); scmc.measureTgHbn_(obs, cfg)
            xi = rng.standard_normal((1, 3, cfg.nSites, 2 * cfg.d)); u = rng.random((1, 3, cfg.nSites))
            cfg.sweep_deterministic([1.0, 2.0, 3.0], [0.3, 0.3, 0.3], xi, u, hits=1)
            scmc.localUpdate_(cfg, 0.0, 0, 1.0, 3, 0.3)
        cfg.close()
        print("ok", model, lat, L, n, prec, "mode", mode, flush=True)
print("sanitize workload done")
