"""Monte Carlo drivers -- host mirror of /root/reference/src/MonteCarlo.jl.  All sweeps run on the device; the host only
forwards schedules and collects results.  Julia's `f!` is spelled `f_`."""
import ctypes as C

import numpy as np

from . import _lib


def _per_replica(x, R):
    return np.broadcast_to(np.asarray(x, dtype=np.float64), (R,)).copy()


def getRandomState(d, rng=None):
    """Haar-random unit vector in C^d (MonteCarlo.jl:244-246)."""
    rng = np.random.default_rng() if rng is None else rng
    v = rng.standard_normal(d) + 1j * rng.standard_normal(d)
    return v / np.linalg.norm(v)


def getRandomState_(newState, state, sigma, rng=None):
    """Proposal normalize(state + sigma * xi) (MonteCarlo.jl:249-260); host utility (the device generates its own proposals)."""
    rng = np.random.default_rng() if rng is None else rng
    d = len(state)
    v = np.asarray(state) + sigma * (rng.standard_normal(d) + 1j * rng.standard_normal(d))
    newState[:] = v / np.linalg.norm(v)


def localUpdate_(cfg, E, accepted_updates, beta, i, sigma, rng=None, replica=0, xi=None, u=None):
    """localUpdate! (MonteCarlo.jl:263-286) of site i on the device with host-drawn (or injected) randoms; returns (E, accepted)."""
    rng = np.random.default_rng() if rng is None else rng
    xi = rng.standard_normal(2 * cfg.d) if xi is None else np.ascontiguousarray(xi, dtype=np.float64)
    u = rng.random() if u is None else float(u)
    dE, acc = C.c_double(), C.c_int32()
    _lib.check(cfg._lib.scmc_update_deterministic(cfg._h, replica, i, _lib.dptr(xi), u, float(beta), float(sigma), C.byref(dE), C.byref(acc)))
    if acc.value:
        E += dE.value
        accepted_updates += 1
    return E, accepted_updates


def run_(cfg, obs, T, T_i, N_therm, N_measure, filename=None, *, measurement_rate=10, checkpoint_rate=100000, report_rate=100,
         current_sweep=0, energy=None, βs=None, σ=60.0, hits=2, keep_series=True, verbose=False):
    """run! (MonteCarlo.jl:3-108) for every replica of cfg at once.  T, T_i, σ may be scalars or per-replica arrays.
    Returns a dict with the final cone widths, acceptance rates, the measurement series [rows, R, 21] and thermalisation log."""
    R = cfg.n_replicas
    T, T_i, sig = _per_replica(T, R), _per_replica(T_i, R), _per_replica(σ, R)
    rate = int(measurement_rate)
    if obs is not None and not hasattr(obs, "push_rows"):
        return _run_with_own_measure(cfg, obs, T, T_i, sig, int(N_therm), int(N_measure), rate, int(current_sweep), int(hits), energy, βs,
                                     filename, verbose)
    n_rows = max(0, (N_therm + N_measure) // rate - max(N_therm, current_sweep) // rate)
    series = np.zeros((max(n_rows, 1), R, _lib.NOBS)) if keep_series else None
    mean = np.zeros((R, _lib.NOBS)); accr = np.zeros(R)
    n_log = max(0, N_therm // rate - current_sweep // rate)
    therm = np.zeros((max(n_log, 1), R))
    p = _lib.RunParams(int(N_therm), int(N_measure), rate, int(hits), 0, C.c_uint64(cfg.seed), C.c_uint64(int(current_sweep)),
                       _lib.dptr(T), _lib.dptr(T_i), _lib.dptr(sig))
    out = _lib.RunResults(_lib.dptr(series) if keep_series else None, 0, _lib.dptr(mean), _lib.dptr(accr), _lib.dptr(therm))
    if verbose and filename:
        print(f"{filename}: Starting thermalization sweeps", flush=True)
    _lib.check(cfg._lib.scmc_run_metropolis(cfg._h, C.byref(p), C.byref(out)))
    cfg.sweep_counter = max(cfg.sweep_counter, N_therm + N_measure)
    rows = int(out.n_rows)
    if keep_series and obs is not None and rows:
        obs.push_rows(series[:rows])
    if energy is not None:
        energy.extend(therm[:n_log, 0].tolist())
        if keep_series:
            energy.extend(series[:rows, 0, 0].tolist())
    if βs is not None:
        n_anneal = int(np.ceil(0.75 * N_therm))
        for s in range(current_sweep + 1, N_therm + 1):
            if s % rate == 0:
                t = T_i[0] * (T[0] / T_i[0]) ** ((s - 1) / (n_anneal - 1)) if (s <= n_anneal and n_anneal > 1) else T[0]
                βs.append(1.0 / t)
        βs.extend([1.0 / T[0]] * rows)
    if verbose and filename:
        print(f"{filename}: Calculation finished", flush=True)
    return dict(sigma=sig, acc_rate=accr, series=series[:rows] if keep_series else None, mean=mean, therm_energy=therm[:n_log], rows=rows)


def _run_with_own_measure(cfg, obs, T, T_i, sig, N_therm, N_measure, rate, current_sweep, hits, energy, βs, filename, verbose):
    """run! for an observables type whose measure! needs more than the device's 21-value rows (ObservablesTgHbn: sublattice magnetisations,
    chiralities; the reference dispatches measure!(obs, cfg, E) on the type, MonteCarlo.jl:82-86).  Thermalisation goes through the driver
    (ramp, sigma adaptation, energy log on the device); the measurement phase runs in intervals of `rate` sweeps at frozen beta / sigma with
    `obs.measure(cfg)` after each.  Same visit numbers as the all-device run, so the chain is the one run_(cfg, None, ...) produces."""
    R = cfg.n_replicas
    n_log = max(0, N_therm // rate - current_sweep // rate)
    therm = np.zeros((max(n_log, 1), R)); mean = np.zeros((R, _lib.NOBS)); accr = np.zeros(R)
    p = _lib.RunParams(N_therm, 0, rate, hits, 0, C.c_uint64(cfg.seed), C.c_uint64(current_sweep), _lib.dptr(T), _lib.dptr(T_i), _lib.dptr(sig))
    out = _lib.RunResults(None, 0, _lib.dptr(mean), _lib.dptr(accr), _lib.dptr(therm))
    if verbose and filename:
        print(f"{filename}: Starting thermalization sweeps", flush=True)
    _lib.check(cfg._lib.scmc_run_metropolis(cfg._h, C.byref(p), C.byref(out)))
    beta = 1.0 / T
    acc_tot = np.zeros(R); att_tot = np.zeros(R)
    rows = 0
    s = max(N_therm, current_sweep)              # sweeps done so far; sweep s + 1 uses visit base s * hits (as scmc_run_metropolis does)
    total = N_therm + N_measure
    while s < total:
        n = min(total, (s // rate + 1) * rate) - s
        cfg.sweep_counter = s
        acc, att = cfg.sweep(n, beta, sig, hits=hits)
        acc_tot += acc; att_tot += att
        s += n
        if s % rate == 0:
            obs.measure(cfg)
            rows += 1
    cfg.sweep_counter = max(cfg.sweep_counter, total)
    accr = np.divide(acc_tot, att_tot, out=np.zeros(R), where=att_tot > 0)
    if energy is not None:
        energy.extend(therm[:n_log, 0].tolist())
    if βs is not None:
        n_anneal = int(np.ceil(0.75 * N_therm))
        for q in range(current_sweep + 1, N_therm + 1):
            if q % rate == 0:
                t = T_i[0] * (T[0] / T_i[0]) ** ((q - 1) / (n_anneal - 1)) if (q <= n_anneal and n_anneal > 1) else T[0]
                βs.append(1.0 / t)
        βs.extend([1.0 / T[0]] * rows)
    if verbose and filename:
        print(f"{filename}: Calculation finished", flush=True)
    return dict(sigma=sig, acc_rate=accr, series=None, mean=None, therm_energy=therm[:n_log], rows=rows)


def runAnnealing_(cfg, T_i, filename=None, *, T_fac=0.98, N_max=10_000_000, N_o=0, N_per_T=10000, min_acc_per_site=100,
                  min_accrate=1e-5, checkpoint_rate=100000, σ_min=0.05, verbose=False, βs=None, energy=None, current_sweep=1,
                  σ=60.0, hits=2):
    """runAnnealing! (MonteCarlo.jl:112-238): cooling loop then N_o optimisation sweeps, every replica independently."""
    R = cfg.n_replicas
    p = _lib.AnnealParams(float(T_i), float(T_fac), int(N_max), int(N_per_T), int(min_acc_per_site), float(min_accrate),
                          float(σ_min), float(σ), int(hits), 0, C.c_uint64(cfg.seed))
    E = np.zeros(R); beta = np.zeros(R); sig = np.zeros(R); sweeps = np.zeros(R, dtype=np.int64)
    out = _lib.AnnealResults(_lib.dptr(E), _lib.dptr(beta), _lib.dptr(sig), _lib.lptr(sweeps))
    _lib.check(cfg._lib.scmc_anneal(cfg._h, C.byref(p), C.byref(out)))
    cfg.sweep_counter = max(cfg.sweep_counter, int(sweeps.max()) + 1)
    E_sa = E.copy()
    if N_o > 0:
        _lib.check(cfg._lib.scmc_optimize(cfg._h, int(N_o), _lib.dptr(E)))
    if energy is not None:
        energy.append(float(E[0] / cfg.nSites))
    if βs is not None:
        βs.append(float(beta[0]))
    if verbose and filename:
        print(f"{filename}: Finished calculation. E/N = {E[0] / cfg.nSites}", flush=True)
    return dict(E=E, E_annealed=E_sa, beta=beta, sigma=sig, sweeps=sweeps)


def localOptimization_(cfg, E=None, i=None, n_sweeps=1):
    """Optimisation sweeps (localOptimization!, MonteCarlo.jl:289-329).  The device moves every site (colour by colour) to the
    minimiser of its local energy; a single-site call is not exposed -- `i` is accepted for signature compatibility."""
    En = np.zeros(cfg.n_replicas)
    _lib.check(cfg._lib.scmc_optimize(cfg._h, int(n_sweeps), _lib.dptr(En)))
    return float(En[0]) if cfg.n_replicas == 1 else En
