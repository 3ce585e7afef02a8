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


SQRT2 = float(np.sqrt(2.0))


def sigma_scale(σ_convention):
    """Factor between a caller's cone width and the one the device uses.  "unit" (default): sigma multiplies N(0,1) normals, as the
    source of getRandomState! reads (MonteCarlo.jl:255-256).  "reference": sigma as the reference's runs print and store it -- its
    randn!(local_rng(), ::Vector{Float64}) delivers normals of variance 1/2 (the notebook's adapted sigma trace is reproduced only with
    that variance: tests/test_oracle_physics.py::test_reference_sigma_trace_is_a_variance_half_convention), so sigma_reference =
    sqrt(2) sigma_unit.  The cap of the adaptation rule (60) is applied in device units either way (at the cap the proposal is a
    uniformly random state)."""
    if σ_convention == "unit":
        return 1.0
    if σ_convention == "reference":
        return SQRT2
    raise ValueError("σ_convention must be 'unit' or 'reference'")


def stream_seed(cfg):
    """Key of the counter-based stream for the current driver call: cfg.seed for the first run of a configuration, a different key for every
    later fresh run (a temperature scan, annealing followed by run!), so that two runs never replay the same (site, replica, visit)
    counters.  Resumed runs (current_sweep > 0) keep their epoch; checkpoints store it."""
    return (int(cfg.seed) + 0x9E3779B97F4A7C15 * int(getattr(cfg, "stream_epoch", 0))) % (1 << 64)


def _new_epoch(cfg, resumed):
    if not resumed:
        cfg.stream_epoch = getattr(cfg, "stream_epoch", -1) + 1 if getattr(cfg, "_driver_runs", 0) else getattr(cfg, "stream_epoch", 0)
    cfg._driver_runs = getattr(cfg, "_driver_runs", 0) + 1


def _beta_schedule(T, T_i, N_therm, sweeps):
    """beta of the given sweeps for every replica: [len(sweeps), R] (thermalisation ramp of MonteCarlo.jl:23-27, then 1 / T)."""
    n_anneal = int(np.ceil(0.75 * N_therm))
    out = np.empty((len(sweeps), len(T)))
    for k, q in enumerate(sweeps):
        t = T_i * (T / T_i) ** ((q - 1) / (n_anneal - 1)) if (q <= n_anneal and n_anneal > 1) else T
        out[k] = 1.0 / t
    return out


def run_(cfg, obs, T, T_i, N_therm, N_measure, filename=None, *, measurement_rate=10, checkpoint_rate=100000, report_rate=100,
         current_sweep=0, energy=None, βs=None, σ=60.0, σ_convention="unit", hits=2, keep_series=True, verbose=False, on_checkpoint=None):
    """run! (MonteCarlo.jl:3-108) for every replica of cfg at once.  T, T_i, σ may be scalars or per-replica arrays.

    The sweeps run on the device in chunks that end at multiples of `checkpoint_rate` (rounded up to a multiple of `measurement_rate`):
    after every chunk `on_checkpoint(sweep, info)` is called (launch_ writes the reference's checkpoint + means files there,
    MonteCarlo.jl:56-59, 88-91), and `energy` / `βs` (the caller's lists, replica 0, MonteCarlo.jl:48-49, 84-85) are up to date.
    Observables with a device row interface (ObservablesGeneric without correlations) are fed from the driver's measurement rows; any
    other observables type (correlations, ObservablesTgHbn) has its `measure(cfg)` called once per measurement interval.
    Returns a dict: final cone widths (in the caller's σ convention), acceptance rates, measurement series [rows, R, 21], thermalisation
    energy log, and `energy_series` / `beta_series` [n, R] (one entry per sigma-adaptation / measurement point of every replica)."""
    R = cfg.n_replicas
    N_therm, N_measure, rate, hits, current_sweep = int(N_therm), int(N_measure), int(measurement_rate), int(hits), int(current_sweep)
    T, T_i = _per_replica(T, R), _per_replica(T_i, R)
    scale = sigma_scale(σ_convention)
    sig = _per_replica(σ, R) / scale
    total = N_therm + N_measure
    own_measure = obs is not None and (not hasattr(obs, "push_rows") or getattr(obs, "with_correlations", False))
    _new_epoch(cfg, resumed=current_sweep > 0)
    seed = stream_seed(cfg)
    cp = max(1, int(checkpoint_rate))
    cp = -(-cp // rate) * rate                                   # a chunk never cuts a sigma-adaptation / measurement interval
    series_parts, therm_parts, e_parts, b_parts = [], [], [], []
    mean_sum = np.zeros((R, _lib.NOBS)); rows_total = 0
    acc_w = np.zeros(R); acc_n = 0
    own_acc = np.zeros(R); own_att = np.zeros(R)
    if verbose and filename:
        print(f"{filename}: Starting thermalization sweeps", flush=True)
    s = current_sweep
    while s < total:
        stop = min(total, (s // cp + 1) * cp)
        # ---- one chunk on the device: sweeps s + 1 .. stop
        therm_end = min(N_therm, stop)
        n_log = max(0, therm_end // rate - s // rate)
        log_sweeps = [q for q in range(s + 1, therm_end + 1) if q % rate == 0]
        meas_lo = max(N_therm, s)
        n_rows = max(0, stop // rate - meas_lo // rate) if stop > N_therm else 0
        therm = np.zeros((max(n_log, 1), R)); mean = np.zeros((R, _lib.NOBS)); accr = np.zeros(R)
        if not own_measure:
            series = np.zeros((max(n_rows, 1), R, _lib.NOBS)) if (keep_series or obs is not None) else None
            p = _lib.RunParams(N_therm, N_measure, rate, hits, 0 if stop >= total else stop, C.c_uint64(seed), C.c_uint64(s),
                               _lib.dptr(T), _lib.dptr(T_i), _lib.dptr(sig))
            out = _lib.RunResults(_lib.dptr(series) if series is not None else None, 0, _lib.dptr(mean), _lib.dptr(accr), _lib.dptr(therm))
            _lib.check(cfg._lib.scmc_run_metropolis(cfg._h, C.byref(p), C.byref(out)))
            rows = int(out.n_rows)
            if rows:
                if obs is not None:
                    obs.push_rows(series[:rows])
                if keep_series:
                    series_parts.append(series[:rows].copy())
                e_meas = series[:rows, :, 0] if series is not None else np.repeat(mean[None, :, 0], rows, axis=0)
                mean_sum += mean * rows
                acc_w += accr * rows; acc_n += rows
        else:
            # thermalisation part of the chunk through the driver, measurement part interval by interval with obs.measure(cfg)
            rows = 0
            e_meas = np.zeros((0, R))
            if s < N_therm:
                p = _lib.RunParams(N_therm, 0, rate, hits, 0 if therm_end >= N_therm else therm_end, C.c_uint64(seed), C.c_uint64(s),
                                   _lib.dptr(T), _lib.dptr(T_i), _lib.dptr(sig))
                out = _lib.RunResults(None, 0, _lib.dptr(mean), _lib.dptr(accr), _lib.dptr(therm))
                _lib.check(cfg._lib.scmc_run_metropolis(cfg._h, C.byref(p), C.byref(out)))
            q = max(therm_end, s)
            beta = 1.0 / T
            es = []
            while q < stop:
                n = min(stop, (q // rate + 1) * rate) - q
                cfg.sweep_counter = q                             # sweep q + 1 uses visit base q * hits, as scmc_run_metropolis does
                acc, att = cfg.sweep(n, beta, sig, hits=hits, seed=seed)
                own_acc += acc; own_att += att
                q += n
                if q % rate == 0:
                    obs.measure(cfg)
                    es.append(cfg.energies() / cfg.nSites)
                    rows += 1
            if es:
                e_meas = np.array(es)
        therm_parts.append(therm[:n_log].copy())
        if n_log:
            e_parts.append(therm[:n_log]); b_parts.append(_beta_schedule(T, T_i, N_therm, log_sweeps))
        if rows:
            e_parts.append(e_meas); b_parts.append(np.repeat((1.0 / T)[None], rows, axis=0))
        if energy is not None:
            energy.extend(therm[:n_log, 0].tolist()); energy.extend(np.asarray(e_meas)[:rows, 0].tolist())
        if βs is not None:
            if n_log:
                βs.extend(b_parts[-2 if rows else -1][:, 0].tolist())
            βs.extend([1.0 / T[0]] * rows)
        rows_total += rows
        s = stop
        cfg.sweep_counter = max(cfg.sweep_counter, s)
        if on_checkpoint is not None and s % cp == 0 and s < total:
            on_checkpoint(s, dict(sigma=sig * scale, energy_series=np.concatenate(e_parts) if e_parts else np.zeros((0, R)),
                                  beta_series=np.concatenate(b_parts) if b_parts else np.zeros((0, R)), rows=rows_total))
        if verbose and filename:
            print(f"{filename}: {s}/{total} sweeps", flush=True)
    if verbose and filename:
        print(f"{filename}: Calculation finished", flush=True)
    if own_measure:
        acc_rate = np.divide(own_acc, own_att, out=np.zeros(R), where=own_att > 0)
        series_out, mean_out = None, None
    else:
        acc_rate = acc_w / acc_n if acc_n else np.zeros(R)
        series_out = (np.concatenate(series_parts) if series_parts else np.zeros((0, R, _lib.NOBS))) if keep_series else None
        mean_out = mean_sum / rows_total if rows_total else mean_sum
    therm_all = np.concatenate(therm_parts) if therm_parts else np.zeros((0, R))
    return dict(sigma=sig * scale, acc_rate=acc_rate, series=series_out, mean=mean_out, therm_energy=therm_all, rows=rows_total,
                energy_series=np.concatenate(e_parts) if e_parts else np.zeros((0, R)),
                beta_series=np.concatenate(b_parts) if b_parts else np.zeros((0, R)), stream_epoch=getattr(cfg, "stream_epoch", 0))


def polishFP64_(cfg, n_sweeps, batch=1024, write_back=True):
    """Local-optimisation sweeps (localOptimization!, MonteCarlo.jl:214-225, 289-329) in FP64 on the states of an FP32 configuration.

    Annealing runs best in FP32 (the production kernels), but an FP32 state is only normalised and an FP32 energy only summed to ~1e-7:
    the annealed ground-state energy of the north star (|E/N - exact| <= 1e-8) needs the last optimisation sweeps, and the energy, in
    FP64.  The FP32 states go through the C ABI (scmc_get_state) into a 64-bit handle of the same model, `batch` replicas at a time,
    are optimised there and (optionally) written back.  Returns the FP64 energies [R]."""
    from .configuration import Configuration
    R = cfg.n_replicas
    E = np.zeros(R)
    for r0 in range(0, R, batch):
        n = min(batch, R - r0)
        c64 = Configuration(cfg.latticename, cfg.L, cfg.lattice.unitcell, cfg.interactions, cfg.onsiteInteraction, cfg.n, n_replicas=n,
                            precision=64, device=cfg.device, replica_offset=cfg.replica_offset + r0, seed=cfg.seed, colour=cfg.colour)
        psi = cfg.get_state(r0, n)
        psi /= np.linalg.norm(psi, axis=2, keepdims=True)          # FP32 states are unit vectors to 1e-7 only
        c64.set_state(psi)
        En = np.zeros(n)
        if n_sweeps > 0:
            _lib.check(c64._lib.scmc_optimize(c64._h, int(n_sweeps), _lib.dptr(En)))
        else:
            En = c64.energies()
        E[r0:r0 + n] = En
        if write_back:
            cfg.set_state(c64.get_state(), r0)
        c64.close()
    return E


def runAnnealing_(cfg, T_i, filename=None, *, T_fac=0.98, N_max=10_000_000, N_o=0, N_per_T=10000, min_acc_per_site=100,
                  min_accrate=1e-5, checkpoint_rate=100000, σ_min=0.05, verbose=False, βs=None, energy=None, current_sweep=1,
                  σ=60.0, σ_convention="unit", hits=2, polish_fp64="auto", log_rate="auto"):
    """runAnnealing! (MonteCarlo.jl:112-238): cooling loop then N_o optimisation sweeps, every replica independently.
    polish_fp64 ("auto": when cfg is FP32 and N_o > 0): the optimisation sweeps and the final energies are done in FP64 (polishFP64_), so that
    an FP32 annealing run reaches the exact ground-state energy to 1e-8 as an FP64 one does.
    log_rate: E/N and beta of every replica are logged every `log_rate` sweeps (the reference pushes them every sweep, MonteCarlo.jl:158-159;
    every replica follows its own cooling schedule on the device, so the series is taken on a fixed grid; "auto" = max(N_per_T, N_max / 128): a log point
    ends a persistent-kernel launch and costs a measurement, so ~100 points keep the overhead of the series below 10 % (C4, 512 restarts + 600 optimisation sweeps: 1.69 s with a point every N_per_T = 100 sweeps, 1.13 s at this grid);
    0 = off) -> `energy_series` / `beta_series` [n, R] in the result; the caller's `energy` / `βs` lists get replica 0's series and the final values."""
    R = cfg.n_replicas
    scale = sigma_scale(σ_convention)
    _new_epoch(cfg, resumed=False)
    if log_rate == "auto":
        log_rate = max(int(N_per_T), -(-int(N_max) // 128))
    log_rate = int(min(max(int(log_rate), 0), 2 ** 31 - 1))
    n_cap = int(N_max) // log_rate + 1 if log_rate else 0
    elog = np.zeros((max(n_cap, 1), R)); blog = np.zeros((max(n_cap, 1), R))
    p = _lib.AnnealParams(float(T_i), float(T_fac), int(N_max), int(N_per_T), int(min_acc_per_site), float(min_accrate),
                          float(σ_min) / scale, float(σ) / scale, int(hits), log_rate, C.c_uint64(stream_seed(cfg)))
    E = np.zeros(R); beta = np.zeros(R); sig = np.zeros(R); sweeps = np.zeros(R, dtype=np.int64)
    out = _lib.AnnealResults(_lib.dptr(E), _lib.dptr(beta), _lib.dptr(sig), _lib.lptr(sweeps), _lib.dptr(elog) if log_rate else None,
                             _lib.dptr(blog) if log_rate else None, n_cap, 0)
    _lib.check(cfg._lib.scmc_anneal(cfg._h, C.byref(p), C.byref(out)))
    n_log = int(out.n_log)
    # the grid stops when the last replica stops; a replica's own series ends at its own last sweep
    keep = [int(min(n_log, sweeps[r] // log_rate)) if log_rate else 0 for r in range(R)]
    cfg.sweep_counter = max(cfg.sweep_counter, int(sweeps.max()) + 1)
    E_sa = E.copy()
    polished = False
    if N_o > 0:
        if polish_fp64 is True or (polish_fp64 == "auto" and cfg.precision == 32):
            # a first pass in FP32 (cheap, moves every site next to its minimiser), the remaining sweeps and the energy in FP64
            n32 = int(N_o) // 2
            if n32:
                _lib.check(cfg._lib.scmc_optimize(cfg._h, n32, _lib.dptr(E)))
            E = polishFP64_(cfg, int(N_o) - n32)
            polished = True
        else:
            _lib.check(cfg._lib.scmc_optimize(cfg._h, int(N_o), _lib.dptr(E)))
    if energy is not None:
        energy.extend(elog[:keep[0], 0].tolist()); energy.append(float(E[0] / cfg.nSites))
    if βs is not None:
        βs.extend(blog[:keep[0], 0].tolist()); βs.append(float(beta[0]))
    if verbose and filename:
        print(f"{filename}: Finished calculation. E/N = {E[0] / cfg.nSites}", flush=True)
    return dict(E=E, E_annealed=E_sa, beta=beta, sigma=sig * scale, sweeps=sweeps, polished_fp64=polished, log_rate=log_rate,
                energy_series=elog[:n_log], beta_series=blog[:n_log], series_length=np.array(keep))


def localOptimization_(cfg, E=None, i=None, n_sweeps=1):
    """Optimisation sweeps (localOptimization!, MonteCarlo.jl:289-329).  The device moves every site (colour by colour) to the
    minimiser of its local energy; a single-site call is not exposed -- `i` is accepted for signature compatibility."""
    En = np.zeros(cfg.n_replicas)
    _lib.check(cfg._lib.scmc_optimize(cfg._h, int(n_sweeps), _lib.dptr(En)))
    return float(En[0]) if cfg.n_replicas == 1 else En
