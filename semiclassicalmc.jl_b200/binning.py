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

    @property
    def count(self):
        return len(self.data)

    def _arr(self):
        return np.stack(self.data) if self.data else np.zeros((0,) + self.shape)

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
