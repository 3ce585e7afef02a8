"""Observables -- host mirror of /root/reference/src/Observables/Observables.jl and ObservablesGeneric.jl.
Reductions over the lattice (energy, magnetisation, correlations, structure factor) run on the device through the C ABI."""
import numpy as np

from . import _lib
from .binning import LogBinner, ErrorPropagator, ReplicaSeries, ReplicaPropagator
from .lattice import getKsInBox  # noqa: F401  (re-export, Observables.jl:91-98)


class Observables:
    pass


def getMagnetization(cfg, replica=0):
    """sum_i T_i / N (Observables.jl:16-18) from the device reduction."""
    return cfg.measure()[replica, 5:21]


def getRs(cfg):
    """All r_i - r_j representatives: positions - (positions[0] + b) for every basis vector b (Observables.jl:44-46)."""
    return np.concatenate([cfg.positions - (cfg.positions[0] + b) for b in cfg.basis], axis=0)


def getProject(cfg):
    """m[i, j] with r_i - r_j = rs[m[i, j]] under periodic boundaries (Observables.jl:49-80), 0-based.
    Uses integer cell coordinates instead of the reference's O(N^3) search; identical map."""
    L, nb = cfg.L, len(cfg.basis)
    cells, bl = cfg.lattice.cells, cfg.basisLabels - 1
    rs = getRs(cfg)
    vecs = np.stack(cfg.unitVectors)
    # candidate slot for (cell difference dn, basis pair): the site with cell dn and basis bi, in block bj  ->  verify numerically
    N = cfg.nSites
    dim = cells.shape[1]
    strides = np.array([L ** (dim - 1 - k) for k in range(dim)])
    proj = np.zeros((N, N), dtype=np.int32)
    basis = np.stack(cfg.basis)
    for i in range(N):
        dn = (cells[i][None, :] - cells) % L                      # [N, dim]
        db = basis[bl[i]][None, :] - basis[bl]                    # [N, dim_space]
        r = dn @ vecs + db
        # rs block b holds positions - positions[0] - basis[b]; look for site (cell dn, basis s) with basis[s] - basis[0] - basis[b] = db
        best = np.full(N, np.iinfo(np.int64).max, dtype=np.int64)
        for b in range(nb):
            for s_ in range(nb):
                ok = np.all(np.abs((basis[s_] - basis[0] - basis[b])[None, :] - db) < 1e-10, axis=1)
                idx = b * N + s_ + nb * (dn @ strides)
                best = np.where(ok & (idx < best), idx, best)     # first match in rs order (the reference uses findfirst)
        assert np.all(best < nb * N) and np.allclose(rs[best], r, atol=1e-9)
        proj[i] = best
    return proj


def getCorrelations(cfg, project, replica=0):
    """C[a, m] = sum_{i,j: project[i,j] = m} T_i^a T_j^a (Observables.jl:22-41), on the device."""
    project = np.ascontiguousarray(project, dtype=np.int32)
    n_rs = len(cfg.basis) * cfg.nSites
    Cm = np.zeros((16, n_rs))
    _lib.check(cfg._lib.scmc_correlations(cfg._h, replica, _lib.iptr(project), n_rs, _lib.dptr(Cm)))
    return Cm


def computeStructureFactorVec(correlation, rs, ks):
    """sf[a, k] = (1/|rs|) sum_n cos(k.r_n) C[a, n] (Observables.jl:116-128) for host-side correlation matrices."""
    rs, ks = np.asarray(rs), np.asarray(ks)
    return (np.asarray(correlation) @ np.cos(rs @ ks.T)) / rs.shape[0]


def computeStructureFactor(correlation, rs, ks, components=None):
    """Observables.jl:101-113; default components = rows 1..15 of the reference (0..14 here)."""
    comps = list(range(np.shape(correlation)[0] - 1)) if components is None else list(components)
    return computeStructureFactorVec(correlation, rs, ks)[comps].sum(axis=0)


def _allowed_momentum_indices(cfg, ks):
    """Integer coordinates m' of ks in the reciprocal basis of the L-supercell (k = sum_i m'_i b_i / L) when every k is an allowed momentum
    of the periodic cluster (two-dimensional lattices), else None."""
    lat = cfg.lattice
    if lat.cells.shape[1] != 2 or ks.shape[1] != 2:
        return None
    A = np.stack(lat.unitcell.lattice_vectors) * lat.L            # supercell vectors L a_i
    mf = ks @ A.T / (2 * np.pi)
    mi = np.rint(mf)
    if not np.allclose(mf, mi, atol=1e-6):
        return None
    return mi.astype(np.int64)


def structureFactorSum(cfg, ks, components=None):
    """S(k) summed over `components` (0-based; default 0..14 = the reference's 1:15) for every replica, [R, nk]: the device
    counterpart of computeStructureFactor (Observables.jl:101-113).  Allowed momenta of a 2-D periodic cluster are summed on the device
    (16 times less data to copy back than structureFactor); other momenta go through structureFactor."""
    comps = list(range(15)) if components is None else [int(c) for c in components]
    ks = _lib.f64(ks)
    mi = _allowed_momentum_indices(cfg, ks)
    if mi is None or len(set(comps)) != len(comps):
        return structureFactor(cfg, ks)[:, comps, :].sum(axis=1)
    mask = 0
    for c in comps:
        mask |= 1 << c
    return _structure_factor_grid(cfg, ks, mi, mask)


class StructureFactorPlan:
    """S(k) summed over `components` on a fixed list of allowed momenta of a 2-D periodic cluster, averaged over measurements ON THE DEVICE:
    `accumulate()` after every measurement interval adds the current S(k) of all replicas to a running sum held by the handle (no copy to the
    host), `mean()` returns [R, nk] and the number of measurements.  The momentum tables are prepared once.  Device counterpart of averaging
    correlations and calling computeStructureFactor at the end (ObservablesGeneric.jl:32-46, 49-72; Observables.jl:101-113)."""

    def __init__(self, cfg, ks, components=None):
        comps = list(range(15)) if components is None else [int(c) for c in components]
        ks = _lib.f64(ks)
        mi = _allowed_momentum_indices(cfg, ks)
        if mi is None or len(set(comps)) != len(comps):
            raise ValueError("StructureFactorPlan: needs allowed momenta of a 2-D periodic cluster and distinct components")
        lat = cfg.lattice
        self.cfg, self.nk, self.L, self.nb = cfg, ks.shape[0], lat.L, len(cfg.basis)
        self.mask = 0
        for c in comps:
            self.mask |= 1 << c
        self.m = np.ascontiguousarray(np.mod(mi, lat.L), dtype=np.int32)
        ang = ks @ np.stack(cfg.basis).astype(float).T
        self.phase = np.ascontiguousarray(np.stack([np.cos(ang), np.sin(ang)], axis=-1))
        self.cell = np.ascontiguousarray(lat.cells, dtype=np.int32)
        self.basis = np.ascontiguousarray(lat.basis_labels - 1, dtype=np.int32)

    def accumulate(self):
        c = self.cfg
        _lib.check(c._lib.scmc_structure_factor_accumulate(c._h, self.L, self.L, self.nb, _lib.iptr(self.cell), _lib.iptr(self.basis), self.nk, _lib.iptr(self.m),
                                                           _lib.dptr(self.phase), float(self.nb * c.nSites), self.mask))

    def mean(self, reset=True):
        """([R, nk] mean over the accumulated measurements, their number)."""
        import ctypes as C
        c = self.cfg
        out = np.zeros((c.n_replicas, self.nk)); cnt = C.c_int64(0)
        _lib.check(c._lib.scmc_structure_factor_fetch(c._h, _lib.dptr(out), C.byref(cnt), 1 if reset else 0))
        return (out / cnt.value if cnt.value else out), int(cnt.value)


def _structure_factor_grid(cfg, ks, mi, mask):
    lat = cfg.lattice
    L, nb = lat.L, len(cfg.basis)
    m = np.ascontiguousarray(np.mod(mi, L), dtype=np.int32)
    delta = np.stack(cfg.basis).astype(float)                 # basis offsets delta_b
    ang = ks @ delta.T                                        # [nk, nb]
    phase = np.ascontiguousarray(np.stack([np.cos(ang), np.sin(ang)], axis=-1))
    cell = np.ascontiguousarray(lat.cells, dtype=np.int32)
    basis = np.ascontiguousarray(lat.basis_labels - 1, dtype=np.int32)
    out = np.empty((cfg.n_replicas, ks.shape[0]) if mask else (cfg.n_replicas, 16, ks.shape[0]))
    _lib.check(cfg._lib.scmc_structure_factor_grid(cfg._h, L, L, nb, _lib.iptr(cell), _lib.iptr(basis), ks.shape[0], _lib.iptr(m), _lib.dptr(phase),
                                                   float(nb * cfg.nSites), mask, _lib.dptr(out)))
    return out


def structureFactor(cfg, ks, method="auto"):
    """Device path: S^a(k) = |sum_i T_i^a e^{ik.r_i}|^2 / |rs| for every replica, [R, 16, nk]; equals
    computeStructureFactorVec(getCorrelations(cfg, getProject(cfg)), getRs(cfg), ks) on allowed momenta.
    method "auto": allowed momenta of a 2-D periodic cluster go through the separable DFT over the cell indices (scmc_structure_factor_grid:
    all momenta of a cell at once), anything else through the direct O(N) sum per momentum; "direct" / "grid" force one of them."""
    ks = _lib.f64(ks)
    mi = _allowed_momentum_indices(cfg, ks) if method in ("auto", "grid") else None
    if method == "grid" and mi is None:
        raise ValueError("structureFactor(method='grid'): the momenta are not all allowed momenta of a 2-D periodic cluster")
    if mi is not None:
        return _structure_factor_grid(cfg, ks, mi, 0)
    pos = _lib.f64(cfg.positions)
    out = np.zeros((cfg.n_replicas, 16, ks.shape[0]))
    _lib.check(cfg._lib.scmc_structure_factor(cfg._h, _lib.dptr(ks), _lib.dptr(pos), ks.shape[0], ks.shape[1],
                                              float(len(cfg.basis) * cfg.nSites), _lib.dptr(out)))
    return out


class ObservablesGeneric(Observables):
    """Accumulators of ObservablesGeneric.jl:3-29, one set per replica."""

    def __init__(self, cfg, correlations=False):
        R = cfg.n_replicas
        # block storage over replicas (binning.ReplicaSeries): a measurement of all replicas is appended in one call; x[r] is replica r's binner
        self.energy = ReplicaPropagator(R, 2)
        self.sigma = ReplicaSeries(R)
        self.tau = ReplicaSeries(R)
        self.sigmatau = ReplicaSeries(R)
        self.magnetizationVec = ReplicaSeries(R, np.zeros(16))
        self.rs = getRs(cfg)
        self.with_correlations = correlations
        self.project = getProject(cfg) if correlations else None
        self.correlations = [LogBinner(np.zeros((16, len(cfg.basis) * cfg.nSites))) for _ in range(R)] if correlations else None

    def measure(self, cfg):
        """measure!(obs, cfg, E) for this observables type (ObservablesGeneric.jl:32-46)."""
        measure_(self, cfg)

    def means(self, cfg, beta, replica=0):
        return means(self, cfg, beta, replica)

    def push_rows(self, rows):
        """rows: [n, R, 21] measurement rows from the device."""
        rows = np.asarray(rows, dtype=np.float64)
        if rows.shape[0] == 0:
            return
        self.energy.push_block(rows[:, :, 0:2]); self.sigma.push_block(rows[:, :, 2]); self.tau.push_block(rows[:, :, 3])
        self.sigmatau.push_block(rows[:, :, 4]); self.magnetizationVec.push_block(rows[:, :, 5:21])


def initializeObservables(obstype, cfg, **kw):
    return obstype(cfg, **kw)


def measure_(obs, cfg, E=None):
    """measure! (ObservablesGeneric.jl:32-46): one measurement of every replica; E is re-summed on the device."""
    obs.push_rows(cfg.measure()[None])
    if obs.with_correlations:
        for r in range(cfg.n_replicas):
            obs.correlations[r].push(getCorrelations(cfg, obs.project, r))


def means(obs, cfg, beta, replica=0):
    """The `means/*` group written by saveMeans! (ObservablesGeneric.jl:49-72) as a dict (mean, error)."""
    N = cfg.nSites
    e = obs.energy[replica]
    c = lambda m: beta * beta * (m[1] - m[0] * m[0]) * N
    dc = lambda m: np.array([-2.0 * beta * beta * m[0] * N, beta * beta * N])
    out = {"energy": (e.means()[0], e.std_errors()[0]), "heat": (e.mean(c), e.std_error(dc))}
    for name, b in (("σ", obs.sigma), ("τ", obs.tau), ("στ", obs.sigmatau), ("magnetizationVec", obs.magnetizationVec)):
        out[name] = (b[replica].mean(), b[replica].std_error())
    if obs.with_correlations:
        if obs.correlations[replica].count == 0:
            raise RuntimeError("means(): no correlation measurement was pushed (run_ measures correlations once per interval when obs.with_correlations is set)")
        out["correlations"] = (obs.correlations[replica].mean(), obs.correlations[replica].std_error())
    return out


# ---- tg-hbn observables (ObservablesTgHbn.jl) ----------------------------------------------------------------------
def getPrefactors(v, basis):
    """Coefficients of v in `basis`, rounded to 5 digits (Observables.jl:83-88)."""
    A = np.array([[np.dot(e1, e2) for e2 in basis] for e1 in basis])
    b = np.array([np.dot(e, v) for e in basis])
    return np.round(np.linalg.solve(A, b), 5)


def getSiteLabelsTriangular120(cfg):
    """Sublattice labels 1, 2, 3 of the 120-degree order on the triangular lattice (ObservablesTgHbn.jl:189-209)."""
    lv = [v / np.linalg.norm(v) for v in cfg.unitVectors]
    basis = [lv[0] + lv[1], 2 * lv[1] - lv[0]]
    labels = np.zeros(cfg.nSites, dtype=np.int64)
    for i in range(cfg.nSites):
        v1 = cfg.positions[i]
        if np.all(getPrefactors(v1, basis) % 1 == 0):
            labels[i] = 1
        elif np.all(getPrefactors(v1 + lv[0], basis) % 1 == 0):
            labels[i] = 2
        else:
            labels[i] = 3
    return labels


def _sublattice_magnetization_native(cfg, site_labels):
    """[R, n_labels, 16] as the device reduction writes it (the 16 components contiguous)."""
    labels = np.unique(site_labels)
    lab0 = np.ascontiguousarray(np.searchsorted(labels, site_labels), dtype=np.int32)
    out = np.zeros((cfg.n_replicas, len(labels), 16))
    _lib.check(cfg._lib.scmc_sublattice_magnetization(cfg._h, _lib.iptr(lab0), len(labels), _lib.dptr(out)))
    return out


def getSublatticeMagnetization(cfg, site_labels):
    """[R, 16, n_labels]: mean T over the sites of every label (ObservablesTgHbn.jl:211-221), device reduction."""
    return np.transpose(_sublattice_magnetization_native(cfg, site_labels), (0, 2, 1))


def getChirality(cfg, tri=None):
    """[R, 3] = (kappa_sigma, kappa_tau, kappa_sigmatau): staggered vector chiralities (ObservablesTgHbn.jl:224-249).  `tri`
    defaults to neighbour slots [1, 5, 4, 2] (1-based) of the tg-hbn bond order, as the reference hard-codes."""
    if tri is None:
        tri = cfg.interactionSites[:, [0, 4, 3, 1]]
    tri = np.ascontiguousarray(tri, dtype=np.int32)
    out = np.zeros((cfg.n_replicas, 3))
    _lib.check(cfg._lib.scmc_chirality(cfg._h, _lib.iptr(tri), _lib.dptr(out)))
    return out


_TGHBN_SETS = {"sd": [4, 8, 12], "dx": [1, 2], "dz": [3], "sx": [5, 6, 9, 10, 13, 14], "sz": [7, 11, 15]}   # 0-based component sets


_TGHBN_KEYS = list(_TGHBN_SETS)
_TGHBN_W = np.zeros((len(_TGHBN_KEYS), 16))                  # W[k, a] = 1 when component a belongs to set k
for _k, _key in enumerate(_TGHBN_KEYS):
    _TGHBN_W[_k, _TGHBN_SETS[_key]] = 1.0


def _tghbn_norms(sub):
    """Norms over the component sets of ObservablesTgHbn.jl:47-98 for all replicas at once: subM<set> = mean over the sublattices of
    |subM[set, l]|, m<set> = |mean over the sublattices of subM[set]|.  sub [R, n_sub, 16] (the device reduction's layout): the sums of
    squares over a set are one GEMM (R n_sub x 16) @ (16 x n_sets) -- measured 2-3x faster than a norm per set at 64 ... 8192 replicas, while
    an einsum / batched-matmul form over the [R, 16, n_sub] layout was up to 2x slower at large R."""
    R, ns, _ = sub.shape
    per_sub = np.sqrt((sub * sub).reshape(R * ns, 16) @ _TGHBN_W.T).reshape(R, ns, -1).mean(axis=1)      # [R, n_sets]
    m = sub.mean(axis=1)                                                                                  # [R, 16]
    tot = np.sqrt((m * m) @ _TGHBN_W.T)                                                                   # [R, n_sets]
    out = {}
    for k, key in enumerate(_TGHBN_KEYS):
        out["subM" + key] = per_sub[:, k]
        out["m" + key] = tot[:, k]
    return out


class ObservablesTgHbn(Observables):
    """Accumulators of ObservablesTgHbn.jl:4-44 (one set per replica): energy, sublattice / total magnetisation norms,
    vector chiralities, each with the (x^2, x^4) pair for its Binder cumulant."""
    NAMES = ["subMsd", "subMsx", "subMsz", "subMdx", "subMdz", "msd", "msx", "msz", "mdx", "mdz",
             "spinChirality", "valleyChirality", "spinValleyChirality"]

    def __init__(self, cfg):
        R = cfg.n_replicas
        self.siteLabels = getSiteLabelsTriangular120(cfg)
        self.energy = ReplicaPropagator(R, 2)
        self.series = {n: ReplicaSeries(R) for n in self.NAMES}
        self.binder = {n: ReplicaPropagator(R, 2) for n in self.NAMES}
        self.with_correlations = False

    def measure(self, cfg):
        """measure!(obs, cfg, E) for this observables type (ObservablesTgHbn.jl:47-98)."""
        measureTgHbn_(self, cfg)

    def means(self, cfg, beta, replica=0):
        return meansTgHbn(self, cfg, beta, replica)

    def values(self, cfg):
        """One measurement of every replica: dict name -> [R] (the quantities pushed by measure!, ObservablesTgHbn.jl:47-98)."""
        kap = getChirality(cfg)
        sub = _sublattice_magnetization_native(cfg, self.siteLabels)        # [R, 3, 16]
        out = {"spinChirality": kap[:, 0], "valleyChirality": kap[:, 1], "spinValleyChirality": kap[:, 2]}
        out.update(_tghbn_norms(sub))
        return out


def measureTgHbn_(obs, cfg):
    """measure!(::ObservablesTgHbn, cfg, E) (ObservablesTgHbn.jl:47-98); E re-summed on the device."""
    row = cfg.measure()
    vals = obs.values(cfg)
    obs.energy.push_block(row[None, :, 0:2])
    for n in obs.NAMES:
        x = np.asarray(vals[n], dtype=np.float64)
        obs.series[n].push_block(x[None])
        obs.binder[n].push_block(np.stack([x ** 2, x ** 4], axis=-1)[None])


def meansTgHbn(obs, cfg, beta, replica=0):
    """The `means/*` group of saveMeans!(::ObservablesTgHbn) (ObservablesTgHbn.jl:101-186): Binder cumulant b = 1 - <x^4>/(3 <x^2>^2)."""
    b = lambda v: 1 - v[1] / (3 * v[0] ** 2)
    db = lambda v: np.array([2 * v[1] / (3 * v[0] ** 3), -1 / (3 * v[0] ** 2)])
    N = cfg.nSites
    e = obs.energy[replica]
    out = {"energy": (e.means()[0], e.std_errors()[0]),
           "heat": (e.mean(lambda m: beta * beta * (m[1] - m[0] ** 2) * N), e.std_error(lambda m: np.array([-2 * beta * beta * m[0] * N, beta * beta * N])))}
    for n in obs.NAMES:
        out[n] = (obs.series[n][replica].mean(), obs.series[n][replica].std_error())
        out["binder_" + n] = (obs.binder[n][replica].mean(b), obs.binder[n][replica].std_error(db))
    return out
