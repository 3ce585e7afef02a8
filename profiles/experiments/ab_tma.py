"""A/B of the streaming psi-gather kernels on the C5 lattice (triangular 128x128, n = 1, FP32): direct gather (mode 3) against the
TMA-pipelined kernel (mode 5).  Usage: python profiles/experiments/ab_tma.py [replicas] [sweeps] [modes...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import scmc_b200 as scmc
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
modes = [int(m) for m in sys.argv[3:]] or [3, 5]
HITS = int(os.environ.get("HITS", "2"))
J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
T = np.tile(np.geomspace(0.02, 2.0, 64), R // 64)
beta, sigma = 1 / T, np.clip(0.6 * np.sqrt(T), 0.02, 60)
for mode in modes:
    cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [J], 128, 1, n_replicas=R, precision=32, seed=1, stream=int(os.environ.get("STREAM", "1")))
    cfg.set_sweep_mode(mode)
    cfg.sweep(2, beta, sigma, hits=HITS, want_counts=False)          # warm-up (thermalises a little, loads the module)
    best = 0.0
    for rep in range(3):
        t0 = time.perf_counter()
        acc, att = cfg.sweep(S, beta, sigma, hits=HITS)
        dt = time.perf_counter() - t0
        best = max(best, att.sum() / dt)
        print(f"mode {mode} rep {rep}: {att.sum() / dt:.4e} updates/s  ({dt * 1e3 / S:.3f} ms per sweep, acceptance {acc.sum() / att.sum():.3f})", flush=True)
    print(f"mode {mode} hits {HITS}: best {best:.4e} updates/s = {1e12 * HITS / best:.2f} ps per visit, E/N mean {np.mean(cfg.energies()) / cfg.nSites:.6f}", flush=True)
    cfg.close()
