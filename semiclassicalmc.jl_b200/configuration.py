"""`Configuration` and its accessors -- host mirror of /root/reference/src/Configuration.jl.

The reference keeps one chain per `Configuration`; here a Configuration owns `n_replicas` independent chains of the same
model (temperature points / restarts / copies) resident on one B200 behind a libscmc_b200 handle.  All reference accessors
act on replica 0 unless a `replica=` argument is given.  Site indices are 0-based (Python), bond labels 1-based like the
reference (label l <-> interactions[l-1]).
"""
import ctypes as C

import numpy as np

from . import _lib
from . import generators as _gen
from .lattice import getLatticePeriodic, colourSites


class Configuration:
    """struct Configuration (Configuration.jl:2-24) + constructor (Configuration.jl:27-77)."""

    def __init__(self, latticename, L, uc, interactions, onsiteInteraction, n, *, n_replicas=1, precision=64, device=0,
                 replica_offset=0, seed=None, colour=None, stream=1, generators=None):
        lat = getLatticePeriodic(uc, L)
        self.latticename, self.L, self.n = latticename, int(L), int(n)
        self.lattice = lat
        self.nSites = lat.n_sites
        self.unitVectors = [np.array(v, dtype=float) for v in uc.lattice_vectors]
        self.basis = [np.array(b, dtype=float) for b in uc.sites]
        self.positions = lat.positions
        self.basisLabels = lat.basis_labels
        self.interactionSites = lat.nbr_site            # [N, z], -1 = no bond
        self.interactionLabels = lat.nbr_label          # [N, z], 1-based
        self.interactions = [np.array(J, dtype=float).reshape(16, 16) for J in interactions]
        self.onsiteInteraction = np.array(onsiteInteraction, dtype=float).reshape(16)
        # generators=: caller-supplied tables [16, d, d] (Hermitian, d = 4 or 6) instead of getGenerators(n) -- e.g. the n = 2 parton
        # representation WITH the Jordan-Wigner sign the reference's f omits; they run on the library's dense-table path
        self.generators = _gen.getGenerators(n) if generators is None else np.asarray(generators, dtype=np.complex128)
        self.generatorsSq = _gen.getGeneratorsSq(n) if generators is None else np.stack([g @ g for g in self.generators])
        self.d = self.generators.shape[1]
        self.n_replicas, self.precision, self.device, self.replica_offset = int(n_replicas), int(precision), int(device), int(replica_offset)
        self.colour = np.ascontiguousarray(colourSites(lat) if colour is None else colour, dtype=np.int32)
        self.sweep_counter = 0          # position in the counter-based random stream
        self.stream_epoch = 0           # fresh driver runs after the first use another stream key (montecarlo.stream_seed)
        self.seed = int(np.random.SeedSequence().entropy % (1 << 63)) if seed is None else int(seed)
        n_labels = len(self.interactions)
        if lat.nbr_label.max() > n_labels:
            raise ValueError("bond label without a coupling matrix")
        lib = _lib.load()
        h = C.c_void_p()
        nbr = np.ascontiguousarray(lat.nbr_site, dtype=np.int32)
        lab0 = np.ascontiguousarray(np.maximum(lat.nbr_label - 1, 0), dtype=np.int32)
        J = _lib.f64(np.stack(self.interactions))
        g, gs = self.generators, self.generatorsSq
        parts = [_lib.f64(g.real), _lib.f64(g.imag), _lib.f64(gs.real), _lib.f64(gs.imag)]
        _lib.check(lib.scmc_create(C.byref(h), self.device, self.precision, self.nSites, self.d, self.n_replicas, self.replica_offset,
                                   nbr.shape[1], _lib.iptr(nbr), _lib.iptr(lab0), n_labels, _lib.dptr(J), _lib.dptr(self.onsiteInteraction),
                                   *[_lib.dptr(x) for x in parts], _lib.iptr(self.colour)))
        self._h, self._lib = h, lib
        # random initial product state (Configuration.jl:62)
        _lib.check(lib.scmc_randomize(self._h, C.c_uint64(self.seed)))
        self.stream = 1
        if stream != 1:
            self.set_stream_version(stream)

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.scmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):                      # Base.length(cfg) (Configuration.jl:80-82)
        return self.nSites

    # ---- device state access
    def set_state(self, psi, replica0=0):
        """psi: [count, N, d] (or [N, d]) complex."""
        psi = np.asarray(psi, dtype=np.complex128)
        if psi.ndim == 2:
            psi = psi[None]
        re, im = _lib.f64(psi.real), _lib.f64(psi.imag)
        _lib.check(self._lib.scmc_set_state(self._h, replica0, psi.shape[0], _lib.dptr(re), _lib.dptr(im)))

    def get_state(self, replica0=0, count=None):
        count = self.n_replicas - replica0 if count is None else count
        re = np.empty((count, self.nSites, self.d)); im = np.empty_like(re)
        _lib.check(self._lib.scmc_get_state(self._h, replica0, count, _lib.dptr(re), _lib.dptr(im)))
        return re + 1j * im

    @property
    def state(self):                        # cfg.state (replica 0): [N, d]
        return self.get_state(0, 1)[0]

    def get_T(self, replica=0, squared=False):
        T = np.empty((self.nSites, 16)); Tsq = np.empty((self.nSites, 16)) if squared else None
        _lib.check(self._lib.scmc_get_T(self._h, replica, _lib.dptr(T), _lib.dptr(Tsq)))
        return (T, Tsq) if squared else T

    def randomize(self, seed=None):
        if seed is not None:
            self.seed = int(seed)
        _lib.check(self._lib.scmc_randomize(self._h, C.c_uint64(self.seed)))

    def energies(self):
        E = np.empty(self.n_replicas)
        _lib.check(self._lib.scmc_energy(self._h, _lib.dptr(E)))
        return E

    def local_field(self, replica=0):
        h = np.empty((self.nSites, 16))
        _lib.check(self._lib.scmc_local_field(self._h, replica, _lib.dptr(h)))
        return h

    def info(self):
        a = np.zeros(8, dtype=np.int64)
        _lib.check(self._lib.scmc_info(self._h, _lib.lptr(a)))
        return dict(n_colours=int(a[0]), generator_table=int(a[1]), device_bytes=int(a[2]), resident=bool(a[3] & 1), resident_psi=bool(a[3] & 2), tma=bool(a[3] & 4), tma_tc=bool(a[3] & 8), cluster=int(a[4]),
                    launches=int(a[5]), n_sites=int(a[6]), n_replicas=int(a[7]))

    def verify_tables(self):
        """Number of violations found by the on-device self-check of the colouring, neighbour tables and bulk-copy tile plan (0 = consistent)."""
        n = np.zeros(1, dtype=np.int64)
        _lib.check(self._lib.scmc_verify_tables(self._h, _lib.lptr(n)))
        return int(n[0])

    def set_sweep_mode(self, mode=0, batch=0):
        """0 auto | 1 streaming T-gather | 2 resident T-gather | 3 streaming psi-gather, direct gather | 4 resident psi-gather |
        5 streaming psi-gather, TMA-pipelined where the class decomposes into bulk copies (include/scmc.h)."""
        _lib.check(self._lib.scmc_set_sweep_mode(self._h, mode, batch))

    def set_stream_version(self, version=1):
        """Random stream of the sweeps (include/scmc.h): 1 = Philox4x32-10 / 24-bit Box-Muller, 2 = Philox4x32-7 / 16-bit Box-Muller."""
        _lib.check(self._lib.scmc_set_stream_version(self._h, int(version)))
        self.stream = int(version)

    def set_lanes(self, lanes=0):
        """0 auto | 1 one stream | 2 two replica lanes on side streams (streaming kernels; chains unchanged)."""
        _lib.check(self._lib.scmc_set_lanes(self._h, lanes))

    def sweep(self, n_sweeps, beta, sigma, hits=2, want_counts=True, seed=None):
        """n_sweeps colour-ordered sweeps of every replica (the 2N-update loops, MonteCarlo.jl:40-44); `seed` overrides the stream key
        (the drivers pass the key of their stream epoch, montecarlo.stream_seed)."""
        beta = np.broadcast_to(np.asarray(beta, dtype=np.float64), (self.n_replicas,)).copy()
        sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (self.n_replicas,)).copy()
        acc = np.zeros(self.n_replicas, dtype=np.int64); att = np.zeros(self.n_replicas, dtype=np.int64)
        _lib.check(self._lib.scmc_sweep(self._h, n_sweeps, hits, _lib.dptr(beta), _lib.dptr(sigma), C.c_uint64(self.seed if seed is None else seed),
                                        C.c_uint64(self.sweep_counter), _lib.lptr(acc) if want_counts else None,
                                        _lib.lptr(att) if want_counts else None))
        self.sweep_counter += n_sweeps
        return acc, att

    def sweep_deterministic(self, beta, sigma, xi, u, hits=1):
        R, N, d = self.n_replicas, self.nSites, self.d
        beta = np.broadcast_to(np.asarray(beta, dtype=np.float64), (R,)).copy()
        sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (R,)).copy()
        xi = _lib.f64(xi, (hits, R, N, 2 * d)); u = _lib.f64(u, (hits, R, N))
        acc = np.zeros((hits, R, N), dtype=np.uint8); dE = np.zeros((hits, R, N))
        _lib.check(self._lib.scmc_sweep_deterministic(self._h, hits, _lib.dptr(beta), _lib.dptr(sigma), _lib.dptr(xi), _lib.dptr(u),
                                                      _lib.bptr(acc), _lib.dptr(dE)))
        return acc, dE

    def measure(self):
        obs = np.empty((self.n_replicas, _lib.NOBS))
        _lib.check(self._lib.scmc_measure(self._h, _lib.dptr(obs)))
        return obs


# ---- accessors with the reference's names (Configuration.jl:80-134) ------------------------------------------------
def getPosition(cfg, i):
    return cfg.positions[i]


def getBasis(cfg):
    return cfg.basis


def getInteractionSites(cfg, i):
    row = cfg.interactionSites[i]
    return row[row >= 0]


def getInteractionLabels(cfg, i):
    return cfg.interactionLabels[i][cfg.interactionSites[i] >= 0]


def getInteraction(cfg, label):
    return cfg.interactions[label - 1]


def getOnsiteInteraction(cfg):
    return cfg.onsiteInteraction


def getGenerators(arg):
    """getGenerators(n::Int) (Generators.jl:2) or getGenerators(cfg) (Configuration.jl:108)."""
    return arg.generators if isinstance(arg, Configuration) else _gen.getGenerators(arg)


def getGeneratorsSq(cfg):
    return cfg.generatorsSq


def getState(cfg, i, replica=0):
    return cfg.get_state(replica, 1)[0, i]


def getSpinExpectation(cfg, i=None, replica=0):
    T = cfg.get_T(replica)
    return T if i is None else T[i]


def getSpinSqExpectation(cfg, i=None, replica=0):
    Tsq = cfg.get_T(replica, squared=True)[1]
    return Tsq if i is None else Tsq[i]


def computeSpinExpectation(state, generators):
    """<state|G^a|state> for a = 1..16 (Configuration.jl:137-163); host utility for single vectors."""
    state = np.asarray(state, dtype=np.complex128)
    return np.einsum("j,ajk,k->a", state.conj(), np.asarray(generators), state).real


def computeSpinExpectation_(newT, state, generators):
    newT[:] = computeSpinExpectation(state, generators)


def exchangeEnergy(T_i, M, T_j):
    """T_i^T M T_j (Configuration.jl:199-218)."""
    return float(np.asarray(T_i) @ np.asarray(M) @ np.asarray(T_j))


def onsiteEnergy(Tsq, M):
    return float(np.dot(Tsq, M))


def getEnergy(cfg, replica=None):
    """getEnergy (Configuration.jl:236-253), evaluated on the device; float for replica 0 by default."""
    E = cfg.energies()
    return float(E[0]) if replica is None else (E if replica == "all" else float(E[replica]))


def getEnergyDifference(cfg, i, newT, newTsq, replica=0):
    """dE for a proposed (newT, newTsq) at site i (Configuration.jl:256-275): device local field + on-site term."""
    T, Tsq = cfg.get_T(replica, squared=True)
    h = cfg.local_field(replica)[i]
    return float(np.dot(cfg.onsiteInteraction, np.asarray(newTsq) - Tsq[i]) + np.dot(np.asarray(newT) - T[i], h))
