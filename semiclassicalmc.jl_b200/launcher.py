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
            replica_offset=0, hits=2, correlations="auto", σ_convention="unit", stream=1):
    """launch! (Launcher.jl:49-90).  `T` may be a list of temperatures: each becomes `replicas_per_T` replicas of one launch.

    With an HDF5 `filename` every chain gets the reference's file (`current_sweep`, `energy`, `βs`, `σ`, `state`, `means/*`, `rs`): written
    every `checkpoint_rate` sweeps during the run (checkpoint! + saveMeans!, MonteCarlo.jl:56-59, 88-91) and once more at the end
    (MonteCarlo.jl:103-104).  overwrite = False (Launcher.jl:78-88): finished chains are not run again, and a launch whose files all stop at
    the same earlier sweep is CONTINUED from there (state, sigma, stream position and series are read back; the counter-based stream makes
    it the same chain an uninterrupted run produces -- observables accumulators restart, as the reference's would after its reseed).
    correlations: "auto" measures getCorrelations at every measure! like the reference (ObservablesGeneric.jl:44) when the N x N project
    map is affordable (N <= 2048; the reference's own getProject is O(N^3)), and says so otherwise."""
    Ts = np.repeat(np.atleast_1d(np.asarray(T, dtype=float)), replicas_per_T)
    R = len(Ts)
    writes = _writes_files(filename)
    files = [resultFilename(filename, Ts[r], r if replicas_per_T > 1 else None, R) for r in range(R)] if writes else []
    total = int(N_therm) + int(N_measure)
    resume = None
    if not overwrite and writes and all(os.path.exists(f) for f in files):
        done = [_io.readCheckpointH5(f) for f in files]
        sweeps = sorted({int(d["current_sweep"]) for d in done})
        if sweeps[0] >= total and all("means" in d for d in done):
            print(f"{files[0]}: Already reached {total} sweeps, terminating.", flush=True)
            return dict(files=files, T=Ts, means=[{k: (g["mean"], g["error"]) for k, g in d["means"].items()} for d in done],
                        sigma=np.array([float(d["σ"]) for d in done]), resumed=True)
        if len(sweeps) == 1 and 0 < sweeps[0] < total:
            resume = done
    cfg = initializeCfg(model, latticename, J, L, n, n_replicas=R, precision=precision, device=device, seed=seed,
                        replica_offset=replica_offset, stream=stream)
    if cfg is None:
        return None
    obstype = obstype or ObservablesGeneric
    kw = {}
    if obstype is ObservablesGeneric:
        kw["correlations"] = (cfg.nSites <= 2048) if correlations == "auto" else bool(correlations)
        if correlations == "auto" and not kw["correlations"]:
            print(f"launch_: {cfg.nSites} sites -- means/correlations not measured (N x N project map); use structureFactor / StructureFactorPlan", flush=True)
    obs = initializeObservables(obstype, cfg, **kw)
    current_sweep, sigma0 = 0, 60.0
    e_prev = [np.zeros(0)] * R; b_prev = [np.zeros(0)] * R
    if resume is not None:
        current_sweep = int(resume[0]["current_sweep"])
        cfg.set_state(np.stack([np.asarray(d["state"]).T for d in resume]))
        cfg.seed = int(resume[0]["seed"]); cfg.stream_epoch = int(resume[0].get("stream_epoch", 0)); cfg._driver_runs = 1
        sigma0 = np.array([float(d["σ"]) for d in resume])
        e_prev = [np.asarray(d["energy"], dtype=float) for d in resume]; b_prev = [np.asarray(d["βs"], dtype=float) for d in resume]
        print(f"{files[0]}: Continuing from sweep {current_sweep}", flush=True)

    def write_files(sweep, info):
        for r in range(R):
            energy = np.concatenate([e_prev[r], info["energy_series"][:, r]])
            betas = np.concatenate([b_prev[r], info["beta_series"][:, r]])
            _io.checkpointH5_(files[r], cfg, sweep, betas, energy, info["sigma"], replica=r)
            if info["rows"] > 0:
                _io.saveMeans_(files[r], obs, cfg, 1.0 / Ts[r], replica=r)

    res = run_(cfg, obs, Ts, T_i, N_therm, N_measure, filename, measurement_rate=measurement_rate, checkpoint_rate=checkpoint_rate,
               report_rate=report_rate, hits=hits, current_sweep=current_sweep, σ=sigma0, σ_convention=σ_convention,
               on_checkpoint=write_files if writes else None)
    res["means"] = [obs.means(cfg, 1.0 / Ts[r], r) for r in range(R)] if res["rows"] > 0 else [{} for _ in range(R)]
    res["cfg"], res["obs"], res["T"] = cfg, obs, Ts
    if writes:                               # checkpoint! + saveMeans! at the end of run! (MonteCarlo.jl:103-104), one HDF5 file per chain
        write_files(total, res)
        res["files"] = files
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
            k = int(res["series_length"][r])      # this restart's E/N and beta series (MonteCarlo.jl:158-159, on a grid of log_rate sweeps) + the final values
            _io.checkpointH5_(fn, cfg, int(res["sweeps"][r]), np.append(res["beta_series"][:k, r], res["beta"][r]),
                              np.append(res["energy_series"][:k, r], res["E"][r] / cfg.nSites), res["sigma"], replica=r)
            res["files"].append(fn)
    return res
