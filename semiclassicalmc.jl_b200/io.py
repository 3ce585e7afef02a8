"""Checkpoint / result files -- host mirror of /root/reference/src/IO.jl (`checkpoint!`, `readCheckpoint`, `readCheckpointSA`)
and of the `means/*` groups written by `saveMeans!`.

The reference stores Julia-`serialize`d blobs inside HDF5 (IO.jl:5-21, 37-49); neither HDF5 nor Julia exist in this image, so
the same logical fields go into a NumPy `.npz` container: `state` (the d x N complex matrix of IO.jl:146-150, per replica),
`current_sweep`, `energy`, `βs`, `σ`, `rs`, `means/<name>/{mean,error}`.  Unlike the reference a resumed run is a
true continuation of the chain: the random stream is counter based, so (seed, sweep counter, state) is a complete checkpoint.

Result files in the reference's own HDF5 layout (one file per temperature / chain, what `getmeans` / `getenergies` and the
notebook cells read) are written by `saveMeans_`, `checkpointH5_`, `saveState_` through the pure-NumPy writer in `hdf5.py`
(spec-built, not opened with libhdf5 here -- see its header); the Julia-`serialize`d `cfg` / `obs` blobs are not reproduced.
Arrays of this mirror carry the reference's index order (`state` d x N, `rs` dim x |rs|, `correlations` 16 x |rs|); HDF5.jl stores
a Julia array with its dimensions reversed, so `_julia` reverses the axes on the way to the file and back.
"""
import os

import numpy as np

from . import hdf5


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


# ---- the reference's HDF5 result files ------------------------------------------------------------------------------
def _julia(x):
    """Array in the reference's (Julia) index order <-> HDF5 dataset of HDF5.jl: axes reversed, memory C-contiguous."""
    a = np.asarray(x)
    return np.ascontiguousarray(a.T) if a.ndim > 1 else a


def getRsMatrix(cfg):
    """vectorToMatrix(obs.rs) (Observables.jl:44-46, IO.jl:146-150): dim x (nBasis N)."""
    return np.concatenate([cfg.positions - (cfg.positions[0] + b) for b in cfg.basis], axis=0).T


def saveMeans_(filename, obs, cfg, β, replica=0):
    """saveMeans! (ObservablesGeneric.jl:49-72, ObservablesTgHbn.jl:101-186): `means/<key>/{mean,error}` and `rs` added to the
    HDF5 file of one chain (`h5open(filename, "cw")`); `replica` selects the chain of a packed launch."""
    tree = {}
    for key, (mean, err) in obs.means(cfg, β, replica).items():
        tree[f"means/{key}/mean"] = _julia(mean)
        tree[f"means/{key}/error"] = _julia(err)
    tree["rs"] = _julia(getRsMatrix(cfg))
    hdf5.update(filename, tree)


def checkpointH5_(filename, cfg, current_sweep, βs, energy, σ, replica=0):
    """checkpoint! (IO.jl:5-21, 37-49) without the Julia-serialize blobs: `current_sweep`, `energy`, `βs`, `σ` as the reference names
    them, the chain's `state` (d x N complex, IO.jl:146-150) in place of `cfg`, and the stream position (`seed`, `sweep_counter`,
    `replica_id`) that makes a resumed run a continuation."""
    psi = cfg.get_state()[replica]                                 # [N, d]
    hdf5.write(filename, {"current_sweep": np.int64(current_sweep), "energy": np.asarray(energy, dtype=np.float64),
                          "βs": np.asarray(βs, dtype=np.float64), "σ": np.float64(np.ravel(σ)[replica] if np.ndim(σ) else σ),
                          "state": _julia(vectorToMatrix(psi)), "seed": np.uint64(cfg.seed), "sweep_counter": np.int64(cfg.sweep_counter),
                          "replica_id": np.int64(cfg.replica_offset + replica), "stream_epoch": np.int64(getattr(cfg, "stream_epoch", 0)),
                          "stream_version": np.int64(getattr(cfg, "stream", 1))})


def saveState_(filename, cfg, replica=0):
    """`f["state"] = vectorToMatrix(cfg.state)` at the end of runAnnealing! (MonteCarlo.jl:232-234)."""
    hdf5.update(filename, {"state": _julia(vectorToMatrix(cfg.get_state()[replica]))})


def readCheckpointH5(filename):
    """readCheckpoint / readCheckpointSA (IO.jl:24-34, 52-61) for the HDF5 files above: dict in the reference's index order."""
    def back(v):
        return {k: back(x) for k, x in v.items()} if isinstance(v, dict) else _julia(v)
    return back(hdf5.read(filename))


def getmeansH5(filename, outfile, Ts):
    """getmeans (IO.jl:64-115), same signature: `filename(T)` names the result file of temperature T; every `means/<key>/{mean,error}`
    is collected into an array with the temperature as LAST (Julia) index and written to `outfile` together with `Ts`.
    Unreadable files are reported and left as zeros, like the reference does."""
    Ts = list(Ts)
    first = readCheckpointH5(filename(Ts[0]))["means"]
    out = {}
    for key, g in first.items():
        s = np.shape(g["mean"])
        out[key + "/mean"] = np.zeros(s + (len(Ts),), dtype=np.asarray(g["mean"]).dtype)
        out[key + "/error"] = np.zeros(s + (len(Ts),), dtype=np.asarray(g["error"]).dtype)
    for i, T in enumerate(Ts):
        try:
            m = readCheckpointH5(filename(T))["means"]
            for key in first:
                out[key + "/mean"][..., i] = m[key]["mean"]
                out[key + "/error"][..., i] = m[key]["error"]
        except Exception:
            print("ERROR: Could not load means from ", filename(T), flush=True)
    tree = {"Ts": np.asarray(Ts, dtype=np.float64)}
    tree.update({k: _julia(v) for k, v in out.items()})
    hdf5.write(outfile, tree)
    return out


def getenergiesH5(filename, outfile, Ts):
    """getenergies (IO.jl:118-143): the `energy` series of every temperature under the dataset name "$(T)"; a missing file gives
    `zeros(1)` (the reference's catch branch, without its undefined-variable defect)."""
    tree = {}
    for T in Ts:
        try:
            tree[repr(float(T))] = np.asarray(hdf5.read(filename(T))["energy"], dtype=np.float64)
        except Exception:
            tree[repr(float(T))] = np.zeros(1)
            print("ERROR: Could not load measurements from ", filename(T), flush=True)
    hdf5.write(outfile, tree)
    return tree
