"""Checkpoint / result files -- host mirror of /root/reference/src/IO.jl (`checkpoint!`, `readCheckpoint`, `readCheckpointSA`)
and of the `means/*` groups written by `saveMeans!`.

The reference stores Julia-`serialize`d blobs inside HDF5 (IO.jl:5-21, 37-49); neither HDF5 nor Julia exist in this image, so
the same logical fields go into a NumPy `.npz` container: `state` (the d x N complex matrix of IO.jl:146-150, per replica),
`current_sweep`, `energy`, `βs`, `σ`, `rs`, `means/<name>/{mean,error}`.  Unlike the reference a resumed run is a
true continuation of the chain: the random stream is counter based, so (seed, sweep counter, state) is a complete checkpoint.
"""
import numpy as np


def vectorToMatrix(state):
    """[N, d] site states -> d x N matrix (IO.jl:146-150)."""
    return np.asarray(state).T.copy()


def checkpoint_(filename, cfg, obs=None, current_sweep=0, βs=(), energy=(), σ=60.0, means=None):
    """checkpoint! (IO.jl:5-21, 37-49): everything needed to continue or post-process a run."""
    psi = cfg.get_state()                                       # [R, N, d]
    fields = {"state": np.transpose(psi, (0, 2, 1)), "current_sweep": np.int64(current_sweep), "energy": np.asarray(energy, dtype=float),
              "βs": np.asarray(βs, dtype=float), "σ": np.asarray(σ, dtype=float), "seed": np.uint64(cfg.seed),
              "sweep_counter": np.int64(cfg.sweep_counter), "replica_offset": np.int64(cfg.replica_offset),
              "positions": cfg.positions, "rs": np.concatenate([cfg.positions - (cfg.positions[0] + b) for b in cfg.basis], axis=0)}
    for name, (mean, err) in (means or {}).items():
        fields[f"means/{name}/mean"] = np.asarray(mean)
        fields[f"means/{name}/error"] = np.asarray(err)
    np.savez(filename, **fields)


def readCheckpoint(filename, cfg=None):
    """readCheckpoint / readCheckpointSA (IO.jl:24-34, 52-61): returns a dict; with `cfg` given, also restores state + stream position."""
    if not str(filename).endswith(".npz"):
        filename = str(filename) + ".npz"
    with np.load(filename) as f:
        out = {k: f[k] for k in f.files}
    if cfg is not None:
        cfg.set_state(np.transpose(out["state"], (0, 2, 1)))
        cfg.seed = int(out["seed"])
        cfg.sweep_counter = int(out["sweep_counter"])
    return out


def getmeans(filenames, name):
    """getmeans (IO.jl:64-115): collect `means/<name>` over result files; unreadable files are skipped like the reference does."""
    means, errors = [], []
    for fn in filenames:
        try:
            d = readCheckpoint(fn)
            means.append(d[f"means/{name}/mean"]); errors.append(d[f"means/{name}/error"])
        except Exception:
            continue
    return np.array(means), np.array(errors)
