"""A/B of the tensor-core mat-vec experiment (SCMC_MMA=1, k_sweep_colour_psi<float,1,false,true>): C5 lattice, streaming psi-gather kernel.
    python profiles/experiments/ab_mma.py            # runs itself twice (the switch is read once per process)
Prints updates/s and, for the same seed, acceptance counts and energies of both arms (the arms differ by rounding only)."""
import os, sys, subprocess, time, json
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)


def arm(dense):
    import scmc_b200 as scmc, torch
    L, R = 128, 2048
    Jm = J
    if dense:
        rng = np.random.default_rng(7); A = rng.uniform(-1, 1, (16, 16)); Jm = 0.5 * (A + A.T); Jm[0, 0] = 0
    cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [Jm], L, 1, n_replicas=R, precision=32, seed=11)
    cfg.set_sweep_mode(3)                                   # streaming psi-gather
    cfg.set_lanes(int(os.environ.get("AB_LANES", "0")))     # 0 auto (two replica lanes), 1 one stream
    T = np.tile(np.geomspace(0.02, 2.0, 64), R // 64); beta = 1 / T; sigma = np.clip(0.6 * np.sqrt(T), 0.02, 60)
    e0 = cfg.energies() / cfg.nSites
    acc, att = cfg.sweep(4, beta, sigma)
    e1 = cfg.energies() / cfg.nSites
    torch.cuda.synchronize(); t = time.perf_counter()
    n = 0
    for _ in range(4):
        cfg.sweep(10, beta, sigma, want_counts=False); n += 10
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n
    print(json.dumps(dict(mma=os.environ.get("SCMC_MMA", "0"), lanes=os.environ.get("AB_LANES", "0"), dense=dense, ms_per_sweep=1e3 * dt, updates_per_s=2 * L * L * R / dt,
                          acc=int(acc.sum()), att=int(att.sum()), e0=float(e0.mean()), e1_mean=float(e1.mean()), e1=[float(x) for x in e1[:4]])), flush=True)
    cfg.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        arm(sys.argv[1] == "dense")
    else:
        for dense in ("diag", "dense"):
            for rep in range(2):
                for m in ("0", "1"):
                    subprocess.run([sys.executable, os.path.abspath(__file__), dense], env=dict(os.environ, SCMC_MMA=m), check=True)
