"""launch! / launchAnnealing! -- host mirror of /root/reference/src/Launcher.jl:49-90, 142-182 (seed, build, run).
The Slurm / launcher-file generators of Launcher.jl are out of scope: a temperature grid is
packed into ONE launch as replicas instead of one job step per temperature."""
import numpy as np

from .models import initializeCfg
from .observables import initializeObservables, ObservablesGeneric, means
from .montecarlo import run_, runAnnealing_


def launch_(filename, model, n, obstype, latticename, L, J, T, T_i, N_therm, N_measure, *, measurement_rate=10,
            checkpoint_rate=100000, report_rate=100000, seed=None, overwrite=True, precision=32, device=0, replicas_per_T=1,
            replica_offset=0, hits=2):
    """launch! (Launcher.jl:49-90).  `T` may be a list of temperatures: each becomes `replicas_per_T` replicas of one launch."""
    Ts = np.repeat(np.atleast_1d(np.asarray(T, dtype=float)), replicas_per_T)
    cfg = initializeCfg(model, latticename, J, L, n, n_replicas=len(Ts), precision=precision, device=device, seed=seed,
                        replica_offset=replica_offset)
    if cfg is None:
        return None
    obs = initializeObservables(obstype or ObservablesGeneric, cfg)
    res = run_(cfg, obs, Ts, T_i, N_therm, N_measure, filename, measurement_rate=measurement_rate, checkpoint_rate=checkpoint_rate,
               report_rate=report_rate, hits=hits)
    res["means"] = [obs.means(cfg, 1.0 / Ts[r], r) for r in range(len(Ts))]      # saveMeans! of the observables type
    res["cfg"], res["obs"], res["T"] = cfg, obs, Ts
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
    return res
