"""ORACLE -- test infrastructure only.  ctypes front end of oracle/scmc_oracle.c plus the
NumPy restatements of the reference's lattice-geometry observables helpers
(`getRs`, `getProject`, `periodicDistance`, `getPrefactors`, `computeStructureFactor`,
/root/reference/src/Observables/Observables.jl:44-113) and of the BinningAnalysis
quantities used by `saveMeans!` (src/Observables/ObservablesGeneric.jl:49-72).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from . import generators as _gen

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


def build(fast=False, force=False):
    target = "_build/liboracle_fast.so" if fast else "_build/liboracle.so"
    path = os.path.join(_HERE, target)
    src = os.path.join(_HERE, "scmc_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        if force and os.path.exists(path):
            os.remove(path)
        subprocess.check_call(["make", "-s", "-C", _HERE, target])
    return path


_libs = {}


def _lib(fast=False):
    if fast not in _libs:
        lib = C.CDLL(build(fast))
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_int] * 4 + [_ip, _ip] + [_dp] * 6
        lib.orc_energy.restype = C.c_double
        lib.orc_energy_difference.restype = C.c_double
        lib.orc_local_optimization.restype = C.c_double
        for name in ("orc_sweeps_random", "orc_updates_injected", "orc_run_metropolis", "orc_run_annealing", "orc_run_annealing_traced", "orc_sweeps_colour_v1", "orc_sweeps_colour"):
            getattr(lib, name).restype = C.c_long
        lib.orc_local_update.restype = C.c_int
        _libs[fast] = lib
    return _libs[fast]


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    """One reference-shaped configuration (single chain, FP64)."""

    def __init__(self, nbr_site, nbr_label, J, K, n, fast=False, generators=None):
        self.lib = _lib(fast)
        nbr_site = np.ascontiguousarray(nbr_site, dtype=np.int32)
        nbr_label = np.ascontiguousarray(nbr_label, dtype=np.int32)  # 1-based labels, 0 = padding
        self.N, self.z = nbr_site.shape
        J = np.ascontiguousarray(J, dtype=np.float64).reshape(-1, 16, 16)
        g = _gen.get_generators(n) if generators is None else np.asarray(generators)
        gs = np.stack([m @ m for m in g])
        self.d = g.shape[1]
        self.gen, self.gensq = g, gs
        self._keep = [np.ascontiguousarray(x) for x in (g.real, g.imag, gs.real, gs.imag)]
        lab0 = np.ascontiguousarray(np.maximum(nbr_label - 1, 0), dtype=np.int32)
        K = np.ascontiguousarray(K, dtype=np.float64)
        self.h = C.c_void_p(self.lib.orc_create(self.N, self.d, self.z, J.shape[0], _i(nbr_site), _i(lab0), _d(J), _d(K),
                                                *[_d(x) for x in self._keep]))
        self.J, self.K, self.nbr_site = J, K, nbr_site

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    def seed(self, seed):
        self.lib.orc_seed(self.h, C.c_uint64(seed))

    def randomize(self):
        self.lib.orc_randomize(self.h)

    def randomize_v1(self, seed, replica=0):
        self.lib.orc_randomize_v1(self.h, C.c_uint64(seed), C.c_uint32(replica))

    def set_state(self, psi):
        psi = np.asarray(psi, dtype=np.complex128).reshape(self.N, self.d)
        re, im = np.ascontiguousarray(psi.real), np.ascontiguousarray(psi.imag)
        self.lib.orc_set_state(self.h, _d(re), _d(im))

    def get_state(self):
        re = np.empty((self.N, self.d)); im = np.empty((self.N, self.d))
        self.lib.orc_get_state(self.h, _d(re), _d(im))
        return re + 1j * im

    def get_T(self):
        T = np.empty((self.N, 16)); Tsq = np.empty((self.N, 16))
        self.lib.orc_get_T(self.h, _d(T), _d(Tsq))
        return T, Tsq

    def energy(self):
        return self.lib.orc_energy(self.h)

    def local_field(self, i):
        h = np.empty(16)
        self.lib.orc_local_field(self.h, C.c_int(i), _d(h))
        return h

    def spin_expectation(self, psi, squared=False):
        psi = np.asarray(psi, dtype=np.complex128)
        re, im = np.ascontiguousarray(psi.real), np.ascontiguousarray(psi.imag)
        out = np.empty(16)
        g = self._keep[2:] if squared else self._keep[:2]
        self.lib.orc_spin_expectation(C.c_int(self.d), _d(re), _d(im), _d(g[0]), _d(g[1]), _d(out))
        return out

    def energy_difference(self, i, newT, newTsq):
        newT = np.ascontiguousarray(newT, dtype=np.float64); newTsq = np.ascontiguousarray(newTsq, dtype=np.float64)
        return self.lib.orc_energy_difference(self.h, C.c_int(i), _d(newT), _d(newTsq))

    def local_update(self, i, xi, u, beta, sigma):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        dE = C.c_double()
        acc = self.lib.orc_local_update(self.h, C.c_int(i), _d(xi), C.c_double(u), C.c_double(beta), C.c_double(sigma), C.byref(dE))
        return bool(acc), dE.value

    def updates_injected(self, sites, xi, u, beta, sigma):
        sites = np.ascontiguousarray(sites, dtype=np.int32)
        n = sites.size
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(n, 2 * self.d)
        u = np.ascontiguousarray(u, dtype=np.float64)
        acc = np.zeros(n, dtype=np.uint8); dE = np.zeros(n)
        b = np.array([beta], dtype=np.float64)
        self.lib.orc_updates_injected(self.h, C.c_long(n), _i(sites), _d(xi), _d(u), _d(b), C.c_double(sigma),
                                      acc.ctypes.data_as(_up), _d(dE))
        return acc, dE

    def sweeps_random(self, n_sweeps, beta, sigma, E=None):
        e = C.c_double(self.energy() if E is None else E)
        acc = self.lib.orc_sweeps_random(self.h, C.c_long(n_sweeps), C.c_double(beta), C.c_double(sigma), C.byref(e))
        return acc, e.value

    def sweeps_colour_v1(self, n_sweeps, hits, colour, beta, sigma, seed, replica=0, sweep0=0, want_mask=False):
        colour = np.ascontiguousarray(colour, dtype=np.int32)
        mask = np.zeros((n_sweeps, hits, self.N), dtype=np.uint8) if want_mask else None
        acc = self.lib.orc_sweeps_colour_v1(self.h, C.c_long(n_sweeps), C.c_int(hits), _i(colour), C.c_int(int(colour.max()) + 1),
                                            C.c_double(beta), C.c_double(sigma), C.c_uint64(seed), C.c_uint32(replica),
                                            C.c_uint64(sweep0), mask.ctypes.data_as(_up) if want_mask else None)
        return (acc, mask) if want_mask else acc

    def sweeps_colour(self, version, n_sweeps, hits, colour, beta, sigma, seed, replica=0, sweep0=0, want_mask=False):
        """Colour-ordered sweeps driven by random stream `version` (1 | 2): the product's chain definition, restated."""
        colour = np.ascontiguousarray(colour, dtype=np.int32)
        mask = np.zeros((n_sweeps, hits, self.N), dtype=np.uint8) if want_mask else None
        acc = self.lib.orc_sweeps_colour(self.h, C.c_int(version), C.c_long(n_sweeps), C.c_int(hits), _i(colour), C.c_int(int(colour.max()) + 1),
                                         C.c_double(beta), C.c_double(sigma), C.c_uint64(seed), C.c_uint32(replica),
                                         C.c_uint64(sweep0), mask.ctypes.data_as(_up) if want_mask else None)
        return (acc, mask) if want_mask else acc

    def stream(self, version, seed, site, replica, visit):
        xi = np.empty(2 * self.d); u = C.c_double()
        self.lib.orc_stream(C.c_int(version), C.c_uint64(seed), C.c_uint32(site), C.c_uint32(replica), C.c_uint64(visit), C.c_int(self.d), _d(xi), C.byref(u))
        return xi, u.value

    def stream_v1(self, seed, site, replica, visit):
        xi = np.empty(2 * self.d); u = C.c_double()
        self.lib.orc_stream_v1(C.c_uint64(seed), C.c_uint32(site), C.c_uint32(replica), C.c_uint64(visit), C.c_int(self.d), _d(xi), C.byref(u))
        return xi, u.value

    def magnetization(self):
        m = np.empty(16)
        self.lib.orc_magnetization(self.h, _d(m))
        return m

    def measure(self, E=None):
        out = np.empty(21)
        self.lib.orc_measure(self.h, C.c_double(self.energy() if E is None else E), _d(out))
        return out

    def correlations(self, project, n_rs):
        project = np.ascontiguousarray(project, dtype=np.int32)
        Cm = np.zeros((16, n_rs))
        self.lib.orc_correlations(self.h, _i(project), C.c_int(n_rs), _d(Cm))
        return Cm

    def set_normal_scale(self, s):
        """Standard deviation of the proposal normals of the oracle's own chain (1 = N(0,1)); for the sigma-convention experiment."""
        self.lib.orc_set_normal_scale(self.h, C.c_double(s))

    def run_metropolis(self, T, T_i, N_therm, N_measure, measurement_rate=10, sigma=60.0, sigma_trace=False):
        if sigma_trace:
            self._sigma_log = np.zeros(N_therm // measurement_rate + 2)
            self.lib.orc_set_sigma_log(self.h, _d(self._sigma_log), C.c_long(self._sigma_log.size))
        rows = N_measure // measurement_rate + 2
        series = np.zeros((rows, 21)); s = C.c_double(sigma); accr = C.c_double()
        nlog = (N_therm + N_measure) // measurement_rate + 2
        elog = np.zeros(nlog); blog = np.zeros(nlog)
        n = self.lib.orc_run_metropolis(self.h, C.c_double(T), C.c_double(T_i), C.c_long(N_therm), C.c_long(N_measure),
                                        C.c_long(measurement_rate), C.byref(s), _d(series), _d(elog), _d(blog), C.byref(accr))
        out = dict(series=series[:n], sigma=s.value, acc_rate=accr.value, energy_log=elog, beta_log=blog)
        if sigma_trace:
            self.lib.orc_set_sigma_log(self.h, None, C.c_long(0))
            out["sigma_trace"] = self._sigma_log[:N_therm // measurement_rate].copy()      # cone width after the adaptation at sweep rate, 2 rate, ...
        return out

    def run_annealing(self, T_i, T_fac=0.98, N_max=10_000_000, N_o=0, N_per_T=10000, min_acc_per_site=100, min_accrate=1e-5,
                      sigma_min=0.05, sigma=60.0, trace_cap=0):
        s = C.c_double(sigma); E = C.c_double(); b = C.c_double()
        trace = np.zeros(max(trace_cap, 1))
        n = self.lib.orc_run_annealing(self.h, C.c_double(T_i), C.c_double(T_fac), C.c_long(N_max), C.c_long(N_o), C.c_long(N_per_T),
                                       C.c_long(min_acc_per_site), C.c_double(min_accrate), C.c_double(sigma_min), C.byref(s),
                                       C.byref(E), C.byref(b), _d(trace) if trace_cap else None, C.c_long(trace_cap))
        return dict(sweeps=n, E=E.value, beta=b.value, sigma=s.value, trace=trace[:min(n, trace_cap)])

    def run_annealing_traced(self, T_i, trace_cap, T_fac=0.98, N_max=10_000_000, N_o=0, N_per_T=10000, min_acc_per_site=100, min_accrate=1e-5,
                             sigma_min=0.05, sigma=60.0):
        """runAnnealing! with the per-sweep series the reference keeps or prints: E/N and beta (MonteCarlo.jl:158-159) and the running
        acceptance rate of the current temperature step (R of the checkpoint line, MonteCarlo.jl:161-167).  Entry k = sweep k + 1."""
        s = C.c_double(sigma); E = C.c_double(); b = C.c_double()
        tr, bt, rt = np.zeros(trace_cap), np.zeros(trace_cap), np.zeros(trace_cap)
        n = self.lib.orc_run_annealing_traced(self.h, C.c_double(T_i), C.c_double(T_fac), C.c_long(N_max), C.c_long(N_o), C.c_long(N_per_T),
                                              C.c_long(min_acc_per_site), C.c_double(min_accrate), C.c_double(sigma_min), C.byref(s),
                                              C.byref(E), C.byref(b), _d(tr), _d(bt), _d(rt), C.c_long(trace_cap))
        k = min(n - 1, trace_cap)
        return dict(sweeps=n, E=E.value, beta=b.value, sigma=s.value, trace=tr[:k], beta_trace=bt[:k - N_o], rate_trace=rt[:k - N_o])

    def local_optimization(self, i):
        return self.lib.orc_local_optimization(self.h, C.c_int(i))


def structure_factor_vec(Cm, rs, ks):
    """computeStructureFactorVec (Observables.jl:116-128) through the C oracle."""
    lib = _lib()
    Cm = np.ascontiguousarray(Cm, dtype=np.float64); rs = np.ascontiguousarray(rs, dtype=np.float64)
    ks = np.ascontiguousarray(ks, dtype=np.float64)
    sf = np.zeros((16, ks.shape[0]))
    lib.orc_structure_factor_vec(_d(Cm), _d(rs), C.c_int(rs.shape[0]), _d(ks), C.c_int(ks.shape[0]), C.c_int(rs.shape[1]), _d(sf))
    return sf


def structure_factor(Cm, rs, ks, components=None):
    """computeStructureFactor (Observables.jl:101-113); default components = rows 1..15 (1-based) = 0..14 here."""
    comps = list(range(Cm.shape[0] - 1)) if components is None else list(components)
    return structure_factor_vec(Cm, rs, ks)[comps].sum(axis=0)


# ---- geometry helpers of Observables.jl, restated ------------------------------------------------------------
def get_rs(positions, basis):
    """getRs (Observables.jl:44-46): positions - (positions[0] + b) for every basis vector b, concatenated."""
    return np.concatenate([positions - (positions[0] + b) for b in basis], axis=0)


def get_prefactors(v, unit_vectors):
    """getPrefactors (Observables.jl:83-88)."""
    A = np.array([[np.dot(e1, e2) for e2 in unit_vectors] for e1 in unit_vectors])
    b = np.array([np.dot(e, v) for e in unit_vectors])
    return np.round(np.linalg.solve(A, b), 5)


def periodic_distance(i, j, positions, basis, basis_labels, unit_vectors, L):
    """periodicDistance (Observables.jl:65-80); basis_labels are 1-based."""
    db = basis[basis_labels[i] - 1] - basis[basis_labels[j] - 1]
    r = positions[i] - positions[j] - db
    ns = get_prefactors(r, unit_vectors)
    ns = ((ns % L) + L) % L
    return sum(n * e for n, e in zip(ns, unit_vectors)) + db


def get_project(positions, basis, basis_labels, unit_vectors, L):
    """getProject (Observables.jl:49-62), 0-based indices; O(N^3) like the reference -- small lattices only."""
    rs = get_rs(positions, basis)
    N = positions.shape[0]
    proj = np.zeros((N, N), dtype=np.int32)
    for i in range(N):
        for j in range(N):
            r = periodic_distance(i, j, positions, basis, basis_labels, unit_vectors, L)
            hit = np.nonzero(np.all(np.abs(rs - r) <= 1e-10, axis=1))[0]
            proj[i, j] = hit[0]
    return proj, rs


# ---- BinningAnalysis restatement (LogBinner mean / std_error at the reliable level) ---------------------------
def log_binning_error(x):
    """Standard error of the mean from logarithmic binning (BinningAnalysis.LogBinner `std_error`):
    bin pairs level by level; use the highest level that still holds >= 32 bins (`_reliable_level`)."""
    x = np.asarray(x, dtype=np.float64)
    errs = []
    lvl = x
    while lvl.shape[0] >= 32:
        n = lvl.shape[0]
        errs.append(np.sqrt(np.var(lvl, axis=0, ddof=1) / n))
        lvl = 0.5 * (lvl[0:n - n % 2:2] + lvl[1:n - n % 2 + 1:2])
    return errs[-1] if errs else np.sqrt(np.var(x, axis=0, ddof=1) / max(x.shape[0], 1))


def specific_heat(e, e2, beta, N):
    """c = beta^2 (<e^2> - <e>^2) N (ObservablesGeneric.jl:63-64)."""
    return beta * beta * (np.mean(e2) - np.mean(e) ** 2) * N


# ---- ObservablesTgHbn.jl restated (NumPy loops, FP64) ---------------------------------------------------------------
def site_labels_triangular120(positions, unit_vectors):
    """getSiteLabelsTriangular120 (src/Observables/ObservablesTgHbn.jl:189-209)."""
    lv = [np.asarray(v) / np.linalg.norm(v) for v in unit_vectors]
    basis = [lv[0] + lv[1], 2 * lv[1] - lv[0]]
    labels = np.zeros(len(positions), dtype=np.int64)
    for i, v1 in enumerate(positions):
        v2 = v1 + lv[0]
        if np.all(get_prefactors(v1, basis) % 1 == 0):
            labels[i] = 1
        elif np.all(get_prefactors(v2, basis) % 1 == 0):
            labels[i] = 2
        else:
            labels[i] = 3
    return labels


def sublattice_magnetization(T, labels):
    """getSublatticeMagnetization (ObservablesTgHbn.jl:211-221): [16, n_labels]."""
    uniq = sorted(set(labels.tolist()))
    m = np.zeros((16, len(uniq)))
    for c, lab in enumerate(uniq):
        spins = [T[j] for j in range(len(T)) if labels[j] == lab]
        m[:, c] = sum(spins) / len(spins)
    return m


def _chirality3(v1, v2, v3):
    """getChirality(v1, v2, v3) (ObservablesTgHbn.jl:246-249)."""
    return 2 / 3 / np.sqrt(3) * (np.cross(v1, v2) + np.cross(v2, v3) + np.cross(v3, v1))[2]


def chirality(T, nbr_site, components):
    """getChirality(cfg, components) (ObservablesTgHbn.jl:224-243); components 0-based; neighbour slots [1,5,4,2] -> 0-based [0,4,3,1]."""
    kap = 0.0
    N = len(T)
    for i in range(N):
        i2, i3, i4, i5 = nbr_site[i][[0, 4, 3, 1]]
        s1, s2, s3, s4, s5 = (T[j][components] for j in (i, i2, i3, i4, i5))
        kap += _chirality3(s1, s2, s3) - _chirality3(s1, s4, s5)
    return kap / N / 2
