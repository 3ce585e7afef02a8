"""CPU tests of the host-side mirror: model tables (Models.jl), colouring, binning, import shim."""
import numpy as np
import pytest


def test_tghbn_tables_follow_reference(scmc):
    """Models.jl:48-113: J/8 scaling, on-site vector, label-2 matrix = transpose of label-1, zeroed [1,1], diagonal nnn."""
    from scmc_b200 import lattice as pl
    captured = {}

    class Fake:
        def __init__(self, latticename, L, uc, interactions, onsite, n, **kw):
            captured.update(name=latticename, L=L, uc=uc, J=interactions, K=onsite, n=n)

    import scmc_b200.models as M
    orig = M.Configuration
    M.Configuration = Fake
    try:
        M.initializeCfgTgHbn([8.0, 1.6, 0.8, 0.4, 4.0], 6, n=2)
    finally:
        M.Configuration = orig
    J1, J2, J3 = captured["J"]
    assert np.allclose(J2, J1.T) and J1[0, 0] == 0 and J2[0, 0] == 0
    assert np.isclose(J1[1, 1], 1.0 + 0.1) and np.isclose(J1[1, 2], 0.05) and np.isclose(J1[2, 1], -0.05) and np.isclose(J1[3, 3], 1.0)
    assert np.allclose(J3, np.diag([0] + [0.2] * 15))
    assert np.allclose(captured["K"], [0, 0, 0, 0.25] + [-0.25, 0, 0, 0.25] * 3)
    uc = captured["uc"]
    assert [b.label for b in uc.bonds] == [1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 3, 3]
    lat = pl.getLatticePeriodic(uc, 6)
    col = pl.colourSites(lat)
    assert col.max() + 1 == 4
    for k in range(12):
        assert np.all(col != col[lat.nbr_site[:, k]])
    # every bond has its reverse with the transposed coupling
    for i in range(lat.n_sites):
        for k in range(12):
            j, l = lat.nbr_site[i, k], lat.nbr_label[i, k]
            back = [lat.nbr_label[j, k2] for k2 in range(12) if lat.nbr_site[j, k2] == i]
            assert any(np.allclose(captured["J"][l - 1], captured["J"][lb - 1].T) for lb in back)


def test_unknown_model_returns_none(scmc, caplog):
    assert scmc.initializeCfg("no-such-model", "triangular", [np.eye(16)], 3, 1) is None      # Models.jl:15


def test_colourings(scmc):
    from scmc_b200 import lattice as pl
    for name, L, ncol in (("triangular", 6, 3), ("triangular", 128, 6), ("triangular", 48, 3), ("triangular", 36, 3), ("honeycomb", 96, 2),
                          ("triangular", 5, None), ("square", 4, 2)):
        lat = pl.getLatticePeriodic(pl.getUnitcell(name), L)
        col = pl.colourSites(lat)
        if ncol:
            assert col.max() + 1 == ncol
        for k in range(lat.nbr_site.shape[1]):
            assert np.all(col != col[lat.nbr_site[:, k]])
    # L % 3 != 0: three large classes + O(L) seam sites in the extra ones (each T is read by 2 other passes instead of 3)
    lat = pl.getLatticePeriodic(pl.getUnitcell("triangular"), 128)
    sizes = np.bincount(pl.colourSites(lat))
    assert sizes[:3].sum() > 0.98 * lat.n_sites and sizes[:3].min() > 5000
    assert pl.colourSites(lat, prefer_few_passes=False).max() + 1 == 4


def test_log_binner_against_oracle_formula(scmc, oracle_mod):
    rng = np.random.default_rng(0)
    x = np.zeros(4096)
    for t in range(1, 4096):
        x[t] = 0.9 * x[t - 1] + rng.standard_normal()          # AR(1): correlated series
    b = scmc.LogBinner()
    b.extend(x)
    assert np.isclose(b.mean(), x.mean())
    assert np.isclose(b.std_error(), oracle_mod.log_binning_error(x))
    naive = x.std(ddof=1) / np.sqrt(len(x))
    assert b.std_error() > 2.5 * naive
    ep = scmc.ErrorPropagator(2)
    for v in x:
        ep.push(v, v * v)
    var = ep.mean(lambda m: m[1] - m[0] ** 2)
    assert np.isclose(var, x.var())
    assert ep.std_error(lambda m: np.array([-2 * m[0], 1.0])) > 0


def test_ks_in_box(scmc):
    ks = scmc.getKsInBox("triangular", 6, [8 * np.pi, 8 * np.pi], [0.0, 0.0])
    assert np.all(np.abs(ks) <= 4 * np.pi + 1e-9)
    from scmc_b200 import lattice as pl
    pos = pl.getLatticePeriodic(pl.getUnitcell("triangular"), 6).positions
    sup = 6 * np.stack(pl.getUnitcell("triangular").lattice_vectors)
    assert np.allclose(np.exp(1j * ks @ sup.T), 1.0)           # allowed momenta of the periodic cluster


def test_replica_block_storage_equals_per_replica_binners(scmc):
    """ObservablesGeneric / ObservablesTgHbn keep their series as blocks over replicas (one append per measurement of all replicas instead of a
    Python push per replica and observable); every replica's view must give exactly the estimators of a binner fed sample by sample."""
    from scmc_b200.binning import LogBinner, ErrorPropagator, ReplicaSeries, ReplicaPropagator
    rng = np.random.default_rng(0)
    R, n = 5, 200
    x = rng.normal(size=(n, R)).cumsum(axis=0) * 0.1 + rng.normal(size=(n, R))
    v = rng.normal(size=(n, R, 16))
    e = rng.normal(size=(n, R, 2))
    S, V, P = ReplicaSeries(R), ReplicaSeries(R, np.zeros(16)), ReplicaPropagator(R, 2)
    for a, b in ((0, 70), (70, 71), (71, 200)):
        S.push_block(x[a:b]); V.push_block(v[a:b]); P.push_block(e[a:b])
    f = lambda m: m[1] - m[0] ** 2
    df = lambda m: np.array([-2 * m[0], 1.0])
    for r in range(R):
        b1, b16, p = LogBinner(), LogBinner(np.zeros(16)), ErrorPropagator(2)
        for t in range(n):
            b1.push(x[t, r]); b16.push(v[t, r]); p.push(e[t, r, 0], e[t, r, 1])
        assert S[r].count == b1.count == n and len(S) == R
        assert S[r].mean() == b1.mean() and S[r].std_error() == b1.std_error() and S[r].reliable_level() == b1.reliable_level()
        assert np.array_equal(V[r].mean(), b16.mean()) and np.array_equal(V[r].std_error(), b16.std_error())
        assert P[r].means() == p.means() and P[r].std_errors() == p.std_errors() and P[r].mean(f) == p.mean(f) and P[r].std_error(df) == p.std_error(df)
    S[2].push(3.0)
    assert S[2].count == n + 1 and S[1].count == n


def test_tghbn_norms_vectorised_equals_reference_formulas(scmc):
    """The component-set norms of ObservablesTgHbn for all replicas in a few array operations vs the formulas of ObservablesTgHbn.jl:47-98
    written out per set (norm over the set's components, mean over the sublattices)."""
    import scmc_b200.observables as ob
    rng = np.random.default_rng(3)
    sub = rng.normal(size=(7, 3, 16))                          # [R, n_sub, 16]: the device reduction's layout
    got = ob._tghbn_norms(sub)
    subM = np.transpose(sub, (0, 2, 1))                        # [R, 16, n_sub] as in the reference's formulas
    m = subM.mean(axis=2)
    for key, comps in ob._TGHBN_SETS.items():
        assert np.allclose(got["subM" + key], np.linalg.norm(subM[:, comps, :], axis=1).mean(axis=1), rtol=1e-13, atol=0)
        assert np.allclose(got["m" + key], np.linalg.norm(m[:, comps], axis=1), rtol=1e-13, atol=0)


def test_3xtf32_split_keeps_fp32_accuracy():
    """Arithmetic of the SCMC_MMA experiment (csrc/scmc_core.cu: matvec_mma) restated in NumPy: operands split into a 19-bit TF32 head
    (mask 0xffffe000) and the exactly representable tail, products a_lo.b_hi + a_hi.b_lo + a_hi.b_hi with the tensor core truncating the
    tails to TF32 as well.  The 16x16 mat-vec then agrees with FP64 to 5e-7 of the sum of magnitudes -- FP32 class, well inside north_star's 1e-5 --
    whereas plain TF32 (heads only) is three orders worse, which is why the kernel issues three MMAs per tile."""
    rng = np.random.default_rng(12)

    def head(x):
        return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    J = rng.uniform(-1, 1, (16, 16)).astype(np.float32)
    P = rng.uniform(-3, 3, (16, 4096)).astype(np.float32)           # density sums of six neighbours
    Jh, Ph = head(J), head(P)
    Jl, Pl = head(J - Jh), head(P - Ph)                             # J - Jh is exact in FP32; the MMA reads its TF32 head
    assert np.array_equal((J - Jh).astype(np.float64), J.astype(np.float64) - Jh.astype(np.float64))
    exact = J.astype(np.float64) @ P.astype(np.float64)
    three = (Jl.astype(np.float64) @ Ph + Jh.astype(np.float64) @ Pl + Jh.astype(np.float64) @ Ph).astype(np.float32)
    one = (Jh.astype(np.float64) @ Ph).astype(np.float32)
    scale = np.abs(J.astype(np.float64)) @ np.abs(P.astype(np.float64))
    err3, err1 = np.max(np.abs(three - exact) / scale), np.max(np.abs(one - exact) / scale)
    assert err3 < 1e-6 and err1 > 1e-4 and err1 < 2e-3
