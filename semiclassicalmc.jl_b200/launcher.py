"""launch! / launchAnnealing! -- host mirror of /root/reference/src/Launcher.jl:49-90, 142-182 (seed, build, run).
The Slurm / launcher-file generators of Launcher.jl are out of scope: a temperature grid is
packed into ONE launch as replicas instead of one job step per temperature."""
import os

import numpy as np

from .models import initializeCfg
from .observables import initializeObservables, ObservablesGeneric, means
from .montecarlo import run_, runAnnealing_
from . import io as _io


def resultFilename(filename, T, replica, n_replicas):
    """File of one chain of a packed launch.  The reference runs one `launch!` per temperature with its own file (Launcher.jl:49-66,
    notebook cell 22: `filename(T)`); here `filename` is either such a function of T, or a path: used as is for a single chain,
    otherwise `<stem>_T<T>[_r<replica>]<.ext>` (replica index only when a temperature occurs more than once)."""
    if callable(filename):
        return filename(T)
    if n_replicas == 1:
        return str(filename)
    stem, dot, ext = str(filename).rpartition(".")
    if not dot:
        stem, ext = ext, "h5"
    return f"{stem}_T{T:.6g}" + (f"_r{replica}" if replica is not None else "") + "." + ext


def _writes_files(filename):
    """HDF5 result files are written for a `filename(T)` function or a path ending in .h5 / .hdf5; any other name is a label for the log."""
    return callable(filename) or (isinstance(filename, str) and filename.lower().endswith((".h5", ".hdf5")))


def launch_(filename, model, n, obstype, latticename, L, J, T, T_i, N_therm, N_measure, *, measurement_rate=10,
            checkpoint_rate=100000, report_rate=100000, seed=None, overwrite=True, precision=32, device=0, replicas_per_T=1,
            replica_offset=0, hits=2):
    """launch! (Launcher.jl:49-90).  `T` may be a list of temperatures: each becomes `replicas_per_T` replicas of one launch."""
    Ts = np.repeat(np.atleast_1d(np.asarray(T, dtype=float)), replicas_per_T)
    if not overwrite and _writes_files(filename):
        # overwrite = false (Launcher.jl:78-88): chains whose files already hold N_therm + N_measure sweeps are not run again.  Files are
        # written once, at the end of a run, so a file is either complete or absent; anything else starts from scratch.
        files = [resultFilename(filename, Ts[r], r if replicas_per_T > 1 else None, len(Ts)) for r in range(len(Ts))]
        if all(os.path.exists(f) for f in files):
            done = [_io.readCheckpointH5(f) for f in files]
            if all(int(d["current_sweep"]) >= N_therm + N_measure and "means" in d for d in done):
                print(f"{files[0]}: Already reached {N_therm + N_measure} sweeps, terminating.", flush=True)
                return dict(files=files, T=Ts, means=[{k: (g["mean"], g["error"]) for k, g in d["means"].items()} for d in done],
                            sigma=np.array([float(d["σ"]) for d in done]), resumed=True)
    cfg = initializeCfg(model, latticename, J, L, n, n_replicas=len(Ts), precision=precision, device=device, seed=seed,
                        replica_offset=replica_offset)
    if cfg is None:
        return None
    obs = initializeObservables(obstype or ObservablesGeneric, cfg)
    res = run_(cfg, obs, Ts, T_i, N_therm, N_measure, filename, measurement_rate=measurement_rate, checkpoint_rate=checkpoint_rate,
               report_rate=report_rate, hits=hits)
    res["means"] = [obs.means(cfg, 1.0 / Ts[r], r) for r in range(len(Ts))]      # saveMeans! of the observables type
    res["cfg"], res["obs"], res["T"] = cfg, obs, Ts
    if _writes_files(filename):              # checkpoint! + saveMeans! at the end of run! (MonteCarlo.jl:103-104), one HDF5 file per chain
        res["files"] = []
        for r in range(len(Ts)):
            fn = resultFilename(filename, Ts[r], r if replicas_per_T > 1 else None, len(Ts))
            series = res.get("series")
            energy = np.concatenate([res["therm_energy"][:, r], series[:, r, 0]]) if series is not None else res["therm_energy"][:, r]
            _io.checkpointH5_(fn, cfg, N_therm + N_measure, [], energy, res["sigma"], replica=r)
            _io.saveMeans_(fn, obs, cfg, 1.0 / Ts[r], replica=r)
            res["files"].append(fn)
    return res


def launchAnnealing_(filename, model, n, latticename, L, J, T_i, *, T_fac=0.98, N_max=10000, N_o=0, N_per_T=100,
                     min_acc_per_site=100, min_accrate=0.0, checkpoint_rate=10000, σ_min=0.05, seed=None, verbose=False,
                     precision=64, device=0, n_restarts=1, replica_offset=0, hits=2):
    """launchAnnealing! (Launcher.jl:142-182) with the launcher's own defaults; `n_restarts` independent restarts."""
    cfg = initializeCfg(model, latticename, J, L, n, n_replicas=n_restarts, precision=precision, device=device, seed=seed,
                        replica_offset=replica_offset)
    if cfg is None:
        return None
    res = runAnnealing_(cfg, T_i, filename, T_fac=T_fac, N_max=N_max, N_o=N_o, N_per_T=N_per_T, min_acc_per_site=min_acc_per_site,
                        min_accrate=min_accrate, checkpoint_rate=checkpoint_rate, σ_min=σ_min, verbose=verbose, hits=hits)
    res["cfg"] = cfg
    res["state"] = cfg.get_state()          # f["state"] (MonteCarlo.jl:232-234), [R, N, d]
    if _writes_files(filename):             # checkpoint! + f["state"] at the end of runAnnealing! (MonteCarlo.jl:230-234), one file per restart
        res["files"] = []
        for r in range(n_restarts):
            stem, _, ext = str(filename).rpartition(".")
            fn = filename(r) if callable(filename) else (str(filename) if n_restarts == 1 else f"{stem}_r{r}.{ext}")
            _io.checkpointH5_(fn, cfg, int(res["sweeps"][r]), [float(res["beta"][r])], [float(res["E"][r] / cfg.nSites)], res["sigma"], replica=r)
            res["files"].append(fn)
    return res
