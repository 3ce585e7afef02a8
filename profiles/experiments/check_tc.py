"""Tensor-core mat-vec variant of the TMA kernel (mode 6) against the FFMA2 form (mode 5) on the C5 lattice: same chain up to rounding.
After one sweep almost every site must hold the same state (a near-tie in an accept test may flip a handful), energies agree to 1e-5."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import scmc_b200 as scmc
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
rng = np.random.default_rng(5)
Jd = rng.normal(size=(16, 16)); Jd = (Jd + Jd.T) / 2; Jd[0, :] = 0; Jd[:, 0] = 0
R = 32
T = np.geomspace(0.02, 2.0, R)
beta, sigma = 1 / T, np.clip(0.6 * np.sqrt(T), 0.02, 60)
for name, JJ in (("notebook J", J), ("dense random J", Jd)):
    for stream in (2, 1):
        out = {}
        for mode in (5, 6):
            cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [JJ], 128, 1, n_replicas=R, precision=32, seed=7, stream=stream)
            cfg.set_sweep_mode(mode)
            acc, att = cfg.sweep(1, beta, sigma, hits=2)
            s1 = cfg.get_state().copy(); e1 = cfg.energies().copy()
            acc2, _ = cfg.sweep(20, beta, sigma, hits=2)
            out[mode] = (acc, s1, e1, acc2, cfg.energies().copy())
            cfg.close()
        a, b = out[5], out[6]
        d = np.abs(a[1] - b[1]).reshape(R, -1, a[1].shape[-1] if a[1].ndim > 2 else 1)
        site_diff = np.abs(a[1] - b[1]).max(axis=-1) if a[1].ndim > 2 else np.abs(a[1] - b[1])
        frac = float((site_diff > 1e-4).mean())
        print(f"{name}, stream {stream}: after 1 sweep accept counts differ by {np.abs(a[0].astype(np.int64) - b[0].astype(np.int64)).max()} of {a[0].mean():.0f},"
              f" entries differing > 1e-4: {frac:.2e}, max |dE|/|E| {np.max(np.abs(a[2] - b[2]) / np.abs(a[2])):.2e};"
              f" after 21 sweeps acceptance {a[3].sum() / (20 * 2 * 16384 * R):.4f} vs {b[3].sum() / (20 * 2 * 16384 * R):.4f},"
              f" mean E/N {a[4].mean() / 16384:.5f} vs {b[4].mean() / 16384:.5f}", flush=True)
        assert frac < 2e-3 and np.max(np.abs(a[2] - b[2]) / np.abs(a[2])) < 1e-4
print("check_tc ok")
