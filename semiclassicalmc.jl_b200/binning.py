"""Host-side error analysis equal in spirit to BinningAnalysis.jl's `LogBinner` / `ErrorPropagator` as used by
`saveMeans!` (/root/reference/src/Observables/ObservablesGeneric.jl:4-9, 49-72).  BinningAnalysis is an unvendored
dependency ("parity unpinned"): this is the textbook logarithmic-binning estimator it implements."""
import numpy as np

MIN_BINS = 32   # BinningAnalysis `_reliable_level`: highest level that still has >= 32 bins


class LogBinner:
    """Accumulates a (vector-valued) time series; mean / std_error at the reliable binning level."""

    def __init__(self, zero=0.0):
        self.shape = np.shape(zero)
        self.data = []

    def push(self, x):
        self.data.append(np.array(x, dtype=np.float64))

    def extend(self, xs):
        for x in xs:
            self.push(x)

    def _arr(self):
        return np.stack(self.data) if self.data else np.zeros((0,) + self.shape)

    @property
    def count(self):
        return self._arr().shape[0]

    def mean(self):
        return self._arr().mean(axis=0)

    def levels(self):
        lvl = self._arr()
        out = []
        while lvl.shape[0] >= 2:
            out.append(lvl)
            n = lvl.shape[0] - lvl.shape[0] % 2
            lvl = 0.5 * (lvl[0:n:2] + lvl[1:n:2])
        return out

    def reliable_level(self):
        lv = self.levels()
        good = [i for i, l in enumerate(lv) if l.shape[0] >= MIN_BINS]
        return good[-1] if good else 0

    def var(self, level=None):
        lv = self.levels()
        if not lv:
            return np.zeros(self.shape)
        l = lv[self.reliable_level() if level is None else level]
        return l.var(axis=0, ddof=1)

    def std_error(self, level=None):
        lv = self.levels()
        if not lv:
            return np.zeros(self.shape)
        l = lv[self.reliable_level() if level is None else level]
        return np.sqrt(l.var(axis=0, ddof=1) / l.shape[0])


class ErrorPropagator:
    """Jointly binned tuple of series with first-order error propagation for f(means) (used for the specific heat)."""

    def __init__(self, n_args=2):
        self.bins = [LogBinner() for _ in range(n_args)]

    def push(self, *xs):
        for b, x in zip(self.bins, xs):
            b.push(x)

    @property
    def count(self):
        return self.bins[0].count

    def means(self):
        return [b.mean() for b in self.bins]

    def std_errors(self):
        return [b.std_error() for b in self.bins]

    def mean(self, f):
        return f(self.means())

    def std_error(self, grad):
        lvls = [b.levels() for b in self.bins]
        if not lvls[0]:
            return 0.0
        lvl = self.bins[0].reliable_level()
        X = np.stack([l[lvl] for l in lvls], axis=1)        # [n_bins, n_args]
        g = np.asarray(grad(self.means()), dtype=np.float64)
        cov = np.atleast_2d(np.cov(X, rowvar=False, ddof=1))
        return float(np.sqrt(abs(g @ cov @ g) / X.shape[0]))


class _ViewBinner(LogBinner):
    """LogBinner of ONE replica over the block storage of a ReplicaSeries (plus anything pushed to it directly)."""

    def __init__(self, parent, pick):
        super().__init__(np.zeros(parent.shape))
        self._parent, self._pick = parent, pick

    def _arr(self):
        base = self._parent._all()
        mine = self._pick(base) if base is not None else np.zeros((0,) + self.shape)
        if self.data:
            return np.concatenate([mine, np.stack(self.data)], axis=0)
        return mine


class ReplicaSeries:
    """Time series of one (vector-valued) observable for R replicas.  Measurements arrive as blocks [n, R, *shape] and are appended in one
    call (no Python loop over replicas per measurement); `series[r]` is that replica's LogBinner (same estimators as a per-replica one)."""

    def __init__(self, R, zero=0.0):
        self.R, self.shape, self.blocks, self._cat, self._views = int(R), np.shape(zero), [], None, {}

    def push_block(self, x):
        x = np.asarray(x, dtype=np.float64)
        assert x.shape[1] == self.R and x.shape[2:] == self.shape, (x.shape, self.R, self.shape)
        self.blocks.append(x.copy())
        self._cat = None

    def _all(self):
        if not self.blocks:
            return None
        if self._cat is None:
            self._cat = self.blocks[0] if len(self.blocks) == 1 else np.concatenate(self.blocks, axis=0)
            self.blocks = [self._cat]
        return self._cat

    def __len__(self):
        return self.R

    def __getitem__(self, r):
        r = range(self.R)[r]
        if r not in self._views:
            self._views[r] = _ViewBinner(self, lambda a, r=r: a[:, r])
        return self._views[r]


class _ViewPropagator(ErrorPropagator):
    def __init__(self, parent, r):
        self.bins = [_ViewBinner(parent._component(i), lambda a, r=r: a[:, r]) for i in range(parent.n_args)]


class _Component:
    """Scalar component i of a ReplicaPropagator's blocks, with the interface _ViewBinner needs."""

    def __init__(self, parent, i):
        self.parent, self.i, self.shape = parent, i, ()

    def _all(self):
        a = self.parent.series._all()
        return None if a is None else a[..., self.i]


class ReplicaPropagator:
    """ErrorPropagator for R replicas: blocks [n, R, n_args]; `prop[r]` is that replica's ErrorPropagator."""

    def __init__(self, R, n_args=2):
        self.R, self.n_args = int(R), int(n_args)
        self.series = ReplicaSeries(R, np.zeros(n_args))
        self._views = {}

    def push_block(self, x):
        self.series.push_block(x)

    def _component(self, i):
        return _Component(self, i)

    def __len__(self):
        return self.R

    def __getitem__(self, r):
        r = range(self.R)[r]
        if r not in self._views:
            self._views[r] = _ViewPropagator(self, r)
        return self._views[r]
