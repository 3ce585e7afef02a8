"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): energies / local fields 1e-12 relative in FP64 and 1e-5 in FP32; accept/reject decisions
bit-exact under injected randoms; thermal observables within statistical error bars; annealed ground-state energy within 1e-8.
"""
import numpy as np
import pytest

from conftest import NOTEBOOK_J, make_oracle, random_states, random_symmetric_J

pytestmark = pytest.mark.gpu

TOL = {64: 1e-12, 32: 1e-5}


def build(scmc, kind, precision, R=3, seed=1234, **kw):
    if kind == "tri_notebook":       # n = 1, triangular L = 6, 3 colours
        return scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 6, 1, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "tri_dense_odd":      # n = 1, dense random symmetric J, L = 5 (greedy colouring)
        return scmc.initializeCfg("nearest-neighbor", "triangular", [random_symmetric_J(7)], 5, 1, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "honey_n2_onsite":    # n = 2 (d = 6) with on-site couplings: Tsq path
        Jon = np.diag(np.linspace(-0.3, 0.4, 16))
        return scmc.initializeCfg("nearest-neighbor", "honeycomb", [Jon, random_symmetric_J(9)], 4, 2, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "tghbn":              # 3 labels, 12 neighbours, on-site K, n = 2, 4 colours
        return scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.15, 0.8], 4, 2, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "tghbn_n1":           # TG/h-BN couplings at filling 1: 3 labels, 12 neighbours, d = 4, no on-site term in the dynamics (several-label psi-gather kernel)
        return scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.15, 0.8], 6, 1, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "honey_n2":           # n = 2 (d = 6), no on-site term (BASELINE config C3 couplings: J = 1_16 with J[1,1] = 0), honeycomb L = 4
        Jc3 = np.eye(16); Jc3[0, 0] = 0.0
        return scmc.initializeCfg("nearest-neighbor", "honeycomb", [Jc3], 4, 2, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "tghbn6":             # TG/h-BN model on triangular L = 6 (3 labels, 12 neighbours, on-site K, n = 2)
        return scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.15, 0.8], 6, 2, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "honey_heis":         # BASELINE config C1: su(4) Heisenberg on honeycomb 6x6
        return scmc.initializeCfg("nearest-neighbor", "honeycomb", [np.eye(16)], 6, 1, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "tri_n3":
        return scmc.initializeCfg("nearest-neighbor", "triangular", [random_symmetric_J(3)], 4, 3, n_replicas=R, precision=precision, seed=seed, **kw)
    if kind == "tri_n3_onsite":      # n = 3 with an on-site matrix: (G^1)^2 = 9 (S(0,0) = 3 1), the constant on-site energy is N sum_a K^a c_a
        return scmc.initializeCfg("nearest-neighbor", "triangular", [np.diag(np.linspace(0.5, -0.4, 16)), random_symmetric_J(3)], 4, 3,
                                  n_replicas=R, precision=precision, seed=seed, **kw)
    raise KeyError(kind)


KINDS = ["tri_notebook", "tri_dense_odd", "honey_n2_onsite", "tghbn", "honey_heis", "tri_n3", "tri_n3_onsite"]


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("kind", KINDS)
def test_expectations_energy_field(scmc, oracle_mod, kind, precision):
    cfg = build(scmc, kind, precision)
    rng = np.random.default_rng(5)
    psi = random_states(rng, cfg.n_replicas, cfg.nSites, cfg.d)
    cfg.set_state(psi)
    tol = TOL[precision]
    back = cfg.get_state()
    assert np.allclose(back, psi, atol=1e-15 if precision == 64 else 1e-7)
    E = cfg.energies()
    for r in range(cfg.n_replicas):
        o = make_oracle(cfg, oracle_mod)
        o.set_state(psi[r])
        T, Tsq = cfg.get_T(r, squared=True)
        To, Tsqo = o.get_T()
        assert np.allclose(T, To, atol=tol * 4) and np.allclose(Tsq, Tsqo, atol=tol * 4)
        Eo = o.energy()
        scale = max(1.0, np.abs(To).sum() * np.abs(np.stack(cfg.interactions)).max())
        assert abs(E[r] - Eo) <= tol * scale, (E[r], Eo)
        h = cfg.local_field(r)
        ho = np.stack([o.local_field(i) for i in range(cfg.nSites)])
        assert np.allclose(h, ho, atol=tol * max(1.0, np.abs(ho).max()) * 4)
    obs = cfg.measure()
    assert np.allclose(obs[:, 0] * cfg.nSites, E, rtol=1e-12, atol=1e-12)
    o = make_oracle(cfg, oracle_mod, replica=0)
    assert np.allclose(obs[0], o.measure(E[0]), atol=10 * tol)


@pytest.mark.parametrize("kind", ["tri_notebook", "honey_n2_onsite", "tghbn"])
def test_single_update_deterministic_fp64(scmc, oracle_mod, kind):
    """localUpdate! with injected randoms: identical decisions, dE to 1e-12, same trajectory."""
    cfg = build(scmc, kind, 64, R=2)
    o = make_oracle(cfg, oracle_mod, replica=1)
    rng = np.random.default_rng(17)
    E = cfg.energies()[1]; acc = 0
    for t in range(150):
        i = int(rng.integers(cfg.nSites))
        xi, u = rng.standard_normal(2 * cfg.d), rng.random()
        beta, sigma = 3.0, 0.4
        ok_o, dE_o = o.local_update(i, xi, u, beta, sigma)
        E_new, acc_new = scmc.localUpdate_(cfg, E, acc, beta, i, sigma, replica=1, xi=xi, u=u)
        assert (acc_new - acc) == int(ok_o)
        if ok_o:
            assert abs((E_new - E) - dE_o) <= 1e-12 * max(1.0, abs(dE_o))
        E, acc = E_new, acc_new
    assert np.allclose(cfg.get_state(1, 1)[0], o.get_state(), atol=1e-12)
    assert abs(cfg.energies()[1] - o.energy()) < 1e-10
    assert abs(E - o.energy()) < 1e-9         # incrementally tracked energy (MonteCarlo.jl:282) stays consistent


@pytest.mark.parametrize("kind", ["tri_notebook", "tri_dense_odd", "honey_n2_onsite", "tghbn"])
def test_sweep_deterministic_bit_exact_masks(scmc, oracle_mod, kind):
    """Colour-ordered sweep with injected randoms: accept masks must equal the oracle's bit for bit (FP64)."""
    R, hits = 3, 2
    cfg = build(scmc, kind, 64, R=R)
    N, d = cfg.nSites, cfg.d
    rng = np.random.default_rng(23)
    xi = rng.standard_normal((hits, R, N, 2 * d)); u = rng.random((hits, R, N))
    beta = np.array([0.5, 2.0, 8.0]); sigma = np.array([1.0, 0.3, 0.1])
    oracles = [make_oracle(cfg, oracle_mod, replica=r) for r in range(R)]
    acc, dE = cfg.sweep_deterministic(beta, sigma, xi, u, hits=hits)
    order = np.argsort(cfg.colour, kind="stable")
    n_ties = 0
    for r, o in enumerate(oracles):
        sites = np.repeat(order, hits)
        hit_idx = np.tile(np.arange(hits), N)
        acc_o, dE_o = o.updates_injected(sites, xi[hit_idx, r, sites], u[hit_idx, r, sites], beta[r], sigma[r])
        p = np.exp(-beta[r] * dE_o)
        ties = np.abs(u[hit_idx, r, sites] - p) < 1e-9
        n_ties += int(ties.sum())
        got_acc, got_dE = acc[hit_idx, r, sites], dE[hit_idx, r, sites]
        assert np.array_equal(got_acc[~ties], acc_o[~ties])
        assert np.allclose(got_dE, dE_o, rtol=1e-12, atol=1e-12)
        assert np.allclose(cfg.get_state(r, 1)[0], o.get_state(), atol=1e-12)
    assert n_ties == 0


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("stream", [1, 2])
@pytest.mark.parametrize("kind", ["tri_notebook", "honey_n2_onsite", "tghbn_n1"])
def test_philox_chain_follows_oracle_fp64(scmc, oracle_mod, kind, stream, mode):
    """Counter-based streams v1 / v2: the FP64 device chain of EVERY sweep kernel (auto, streaming / resident T-gather, streaming /
    resident psi-gather) equals the oracle's restatement of the same colour-ordered chain, step results included."""
    R, off, seed = 3, 40, 987654321
    cfg = build(scmc, kind, 64, R=R, seed=seed, replica_offset=off, stream=stream)
    info = cfg.info()
    if (mode == 2 and not info["resident"]) or (mode == 4 and not info["resident_psi"]):
        pytest.skip("resident kernel of this family not available for the model")
    cfg.set_sweep_mode(mode)
    for r in range(R):
        o = make_oracle(cfg, oracle_mod)
        o.randomize_v1(seed, off + r)
        assert np.allclose(cfg.get_state(r, 1)[0], o.get_state(), atol=1e-13)
    beta, sigma = np.array([1.0, 4.0, 20.0]), np.array([0.8, 0.3, 0.1])
    acc1, att1 = cfg.sweep(3, beta, sigma, hits=2)
    acc2, _ = cfg.sweep(2, beta, sigma, hits=2)          # continues the stream at sweep_counter = 3
    assert np.all(att1 == 3 * 2 * cfg.nSites)
    for r in range(R):
        o = make_oracle(cfg, oracle_mod)
        o.randomize_v1(seed, off + r)
        a1 = o.sweeps_colour(stream, 3, 2, cfg.colour, beta[r], sigma[r], seed, off + r, 0)
        a2 = o.sweeps_colour(stream, 2, 2, cfg.colour, beta[r], sigma[r], seed, off + r, 3)
        assert (acc1[r], acc2[r]) == (a1, a2)
        assert np.allclose(cfg.get_state(r, 1)[0], o.get_state(), atol=1e-10)


def test_replica_sharding_is_invisible(scmc):
    """Same global replica ids on differently sharded handles give bit-identical chains (multi-GPU contract)."""
    seed = 4242
    full = build(scmc, "tri_notebook", 32, R=6, seed=seed)
    lo = build(scmc, "tri_notebook", 32, R=2, seed=seed, replica_offset=0)
    hi = build(scmc, "tri_notebook", 32, R=4, seed=seed, replica_offset=2)
    betas = np.linspace(0.5, 30, 6); sig = np.full(6, 0.3)
    a = full.sweep(4, betas, sig)[0]
    b = np.concatenate([lo.sweep(4, betas[:2], sig[:2])[0], hi.sweep(4, betas[2:], sig[2:])[0]])
    assert np.array_equal(a, b)
    assert np.array_equal(full.get_state(), np.concatenate([lo.get_state(), hi.get_state()]))
    assert np.array_equal(full.energies(), np.concatenate([lo.energies(), hi.energies()]))


@pytest.mark.parametrize("precision", [32, 64])
def test_invariants_after_sweeps(scmc, precision):
    """Size-independent properties on a large lattice (128 x 128 triangular, BASELINE C5 shape, fewer replicas)."""
    cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 128, 1, n_replicas=4, precision=precision, seed=99)
    assert cfg.info()["n_colours"] == 6          # 3 large classes + seam classes (lattice.colourSites)
    E0 = cfg.energies()
    acc, att = cfg.sweep(3, np.array([0.5, 2.0, 10.0, 1e9]), np.array([0.5, 0.3, 0.2, 0.2]))
    assert np.all(acc > 0) and np.all(acc <= att)
    tol = 1e-12 if precision == 64 else 2e-6
    psi = cfg.get_state()
    assert np.allclose(np.linalg.norm(psi, axis=-1), 1.0, atol=tol)
    for r in (0, 3):
        T = cfg.get_T(r)
        assert np.allclose(T[:, 0], 1.0, atol=4 * tol) and np.allclose((T[:, 1:] ** 2).sum(axis=1), 3.0, atol=16 * tol)
        assert np.allclose(T, np.einsum("ij,ajk,ik->ia", psi[r].conj(), cfg.generators, psi[r]).real, atol=8 * tol)
    E1 = cfg.energies()
    assert E1[3] < E0[3]                      # beta = 1e9: only downhill moves are accepted
    acc_cold, _ = cfg.sweep(1, np.full(4, 1e9), np.full(4, 0.2))
    assert acc_cold[3] > 0 and cfg.energies()[3] <= E1[3] + 1e-6 * abs(E1[3])


@pytest.mark.parametrize("stream", [1, 2])
def test_thermal_statistics_match_reference_chain(scmc, oracle_mod, stream):
    """BASELINE config C1 (su(4) Heisenberg, honeycomb 6x6, T = 0.1): FP32 colour-ordered device chains (random stream v1 and v2) vs
    the oracle's sequential random-site chain.  Different chains, same stationary distribution: e and c agree within error bars."""
    R = 48
    cfg = build(scmc, "honey_heis", 32, R=R, seed=31337, stream=stream)
    res = scmc.run_(cfg, None, 0.1, 1.0, 2000, 6000, measurement_rate=10)
    ser = res["series"]                                      # [rows, R, 21]
    e_rep = ser[:, :, 0].mean(axis=0)
    c_rep = 100.0 * (ser[:, :, 1].mean(axis=0) - e_rep ** 2) * cfg.nSites
    e_gpu, e_gpu_err = e_rep.mean(), e_rep.std(ddof=1) / np.sqrt(R)
    c_gpu, c_gpu_err = c_rep.mean(), c_rep.std(ddof=1) / np.sqrt(R)
    es, cs = [], []
    for k in range(6):
        o = make_oracle(cfg, oracle_mod)
        o.seed(1000 + k); o.randomize()
        s = o.run_metropolis(0.1, 1.0, 2000, 6000, 10)["series"]
        es.append(s[:, 0].mean()); cs.append(100.0 * (s[:, 1].mean() - s[:, 0].mean() ** 2) * cfg.nSites)
    e_ref, e_ref_err = np.mean(es), np.std(es, ddof=1) / np.sqrt(len(es))
    c_ref, c_ref_err = np.mean(cs), np.std(cs, ddof=1) / np.sqrt(len(cs))
    assert abs(e_gpu - e_ref) < 4.5 * np.hypot(e_gpu_err, e_ref_err) + 1e-4, (e_gpu, e_gpu_err, e_ref, e_ref_err)
    assert abs(c_gpu - c_ref) < 4.5 * np.hypot(c_gpu_err, c_ref_err) + 0.02, (c_gpu, c_gpu_err, c_ref, c_ref_err)
    assert np.all(np.abs(res["acc_rate"] - 0.5) < 0.1)      # sigma adapted towards 50 % (MonteCarlo.jl:47-54)


@pytest.mark.parametrize("kind,T,stream", [("honey_n2", 0.15, 1), ("honey_n2_onsite", 0.15, 2), ("tghbn6", 0.05, 1), ("tghbn6", 0.05, 2)])
def test_fp32_production_statistics_n2_and_tghbn_match_sequential_chain(scmc, oracle_mod, kind, T, stream):
    """The FP32 production kernels for d = 6 (n = 2: T-gather without, psi-gather with on-site term) and for the TG/h-BN model (three
    coupling labels, 12 neighbours, label-major gather) against the oracle's sequential random-site FP64 chain (the reference's sweep,
    src/MonteCarlo.jl:40-44): energy, specific heat and |m| agree within error bars."""
    R = 48
    cfg = build(scmc, kind, 32, R=R, seed=4711, stream=stream)
    N = cfg.nSites
    res = scmc.run_(cfg, None, T, 1.0, 2000, 6000, measurement_rate=10)
    ser = res["series"]                                      # [rows, R, 21]
    e_rep = ser[:, :, 0].mean(axis=0)
    c_rep = (ser[:, :, 1].mean(axis=0) - e_rep ** 2) * N / T ** 2
    m_rep = ser[:, :, 2:5].mean(axis=0)
    es, cs, ms = [], [], []
    for k in range(6):
        o = make_oracle(cfg, oracle_mod)
        o.seed(2000 + k); o.randomize()
        s_ = o.run_metropolis(T, 1.0, 2000, 6000, 10)["series"]
        es.append(s_[:, 0].mean()); cs.append((s_[:, 1].mean() - s_[:, 0].mean() ** 2) * N / T ** 2); ms.append(s_[:, 2:5].mean(axis=0))
    def cmp(gpu, ref, floor):
        g, ge = np.mean(gpu, axis=0), np.std(gpu, axis=0, ddof=1) / np.sqrt(len(gpu))
        r, re_ = np.mean(ref, axis=0), np.std(ref, axis=0, ddof=1) / np.sqrt(len(ref))
        assert np.all(np.abs(g - r) < 4.5 * np.hypot(ge, re_) + floor), (kind, g, ge, r, re_)
    cmp(e_rep, es, 1e-4 * max(1.0, abs(np.mean(es))))
    cmp(c_rep, cs, 0.03)
    cmp(m_rep, np.array(ms), 2e-3)
    assert np.all(np.abs(res["acc_rate"] - 0.5) < 0.12)


@pytest.mark.parametrize("stream", [1, 2])
def test_notebook_observables_on_device(scmc, golden, stream):
    """doc/examples.ipynb cell 13 (triangular L = 6, T = 0.02): e = -3.540997(91), |m_sigma| = 0.993221(8), |m_tau| = 0.02001(13)."""
    R = 32
    cfg = build(scmc, "tri_notebook", 32, R=R, seed=2023, stream=stream)
    res = scmc.run_(cfg, None, 0.02, 1.0, 4000, 8000, measurement_rate=10)
    m = res["mean"]
    e = m[:, 0]
    g = golden["metropolis"]["means"]
    assert abs(e.mean() - g["energy/mean"]) < 5 * np.hypot(e.std(ddof=1) / np.sqrt(R), g["energy/error"]) + 3e-4, (e.mean(), e.std())
    assert abs(m[:, 2].mean() - g["σ/mean"]) < 1e-4
    assert abs(m[:, 3].mean() - g["τ/mean"]) < 1.5e-3
    assert abs(m[:, 4].mean() - g["στ/mean"]) < 1.5e-3
    c = 2500.0 * (m[:, 1] - m[:, 0] ** 2) * 36
    assert abs(c.mean() - g["heat/mean"]) < 5 * np.hypot(c.std(ddof=1) / np.sqrt(R), g["heat/error"]) + 0.1


def test_annealing_reaches_exact_ground_state(scmc):
    """doc/examples.ipynb cell 25 parameters (lattice L = 6 here): annealing + optimisation sweeps reach E/N = -3.6 within 1e-8 (FP64)."""
    cfg = build(scmc, "tri_notebook", 64, R=8, seed=77)
    res = scmc.runAnnealing_(cfg, 3.0, T_fac=0.98, N_max=20000, N_o=400, N_per_T=100, min_acc_per_site=10, min_accrate=1e-3, σ_min=0.05)
    e_sa, e = res["E_annealed"] / 36, res["E"] / 36
    assert np.all(e <= e_sa + 1e-12)
    assert np.all(e >= -3.6 - 1e-9)
    assert abs(e.min() - (-3.6)) < 1e-8, e
    assert np.all(res["sweeps"] > 100) and np.all(res["sigma"] >= 0.05 - 1e-12)


@pytest.mark.parametrize("kind", ["tri_notebook", "honey_heis"])
def test_correlations_and_structure_factor(scmc, oracle_mod, kind):
    cfg = build(scmc, kind, 64, R=2)
    cfg.sweep(2, 1.0, 0.5)
    proj = scmc.getProject(cfg)
    rs = scmc.getRs(cfg)
    proj_o, rs_o = oracle_mod.get_project(cfg.positions, cfg.basis, cfg.basisLabels, cfg.unitVectors, cfg.L)
    assert np.array_equal(proj, proj_o) and np.allclose(rs, rs_o)
    o = make_oracle(cfg, oracle_mod, replica=1)
    Cg = scmc.getCorrelations(cfg, proj, replica=1)
    Co = o.correlations(proj, rs.shape[0])
    assert np.allclose(Cg, Co, rtol=1e-12, atol=1e-12)
    B = 2 * np.pi * np.linalg.inv(np.stack(cfg.unitVectors) * cfg.L).T
    ks = np.array([a * B[0] + b * B[1] for a in range(-cfg.L, cfg.L + 1) for b in range(-cfg.L, cfg.L + 1)])
    Sg = scmc.structureFactor(cfg, ks)[1]
    So = oracle_mod.structure_factor_vec(Co, rs, ks)
    assert np.allclose(Sg, So, atol=1e-10)
    assert np.allclose(scmc.computeStructureFactor(Cg, rs, ks), oracle_mod.structure_factor(Co, rs, ks), atol=1e-10)


def test_create_rejects_bad_tables(scmc):
    from scmc_b200 import lattice as pl
    uc = pl.getUnitcell("triangular")
    with pytest.raises(scmc._lib.ScmcError, match="colour"):
        scmc.Configuration("triangular", 4, uc, [NOTEBOOK_J], np.zeros(16), 1, colour=np.zeros(16, dtype=np.int32))
    Jasym = NOTEBOOK_J.copy(); Jasym[1, 2] = 0.5
    with pytest.raises(scmc._lib.ScmcError, match="transposed"):
        scmc.Configuration("triangular", 4, uc, [Jasym], np.zeros(16), 1)


@pytest.mark.parametrize("case", [
    ("triangular", 6, 1, 32, 5),       # one CTA per replica
    ("honeycomb", 6, 2, 64, 3),        # n = 2, FP64, one CTA (T-gather family only: d = 6)
    ("triangular", 48, 1, 32, 3),      # 2304 sites
    ("triangular", 48, 1, 64, 2),      # FP64
    ("triangular", 128, 1, 32, 3),     # BASELINE C5 lattice -> psi family: cluster of 4 x 177 KB
])
def test_resident_kernels_equal_streaming_kernels(scmc, case, monkeypatch):
    """Within one arithmetic family the replica-resident cluster kernel and the streaming colour-pass kernel realise the same
    chain: identical accept counts and bit-identical states / energies (same stream v1, colour order and arithmetic).  The two
    families (T-gather: modes 1/2, psi-gather: modes 3/4) agree with each other to rounding."""
    lattice, L, n, precision, R = case
    J = [NOTEBOOK_J] if n == 1 else [np.diag(np.linspace(-0.2, 0.3, 16)), random_symmetric_J(11)]
    mk = lambda: scmc.initializeCfg("nearest-neighbor", lattice, J, L, n, n_replicas=R, precision=precision, seed=555, replica_offset=9)
    beta = np.geomspace(0.5, 40.0, R); sigma = np.geomspace(0.6, 0.08, R)
    finals = {}
    probe = mk(); info = probe.info(); probe.close()
    if (lattice, L, precision) == ("triangular", 128, 32):
        # by size this lattice needs 4 CTAs x 177 KB; with 3 replicas scmc_create grows the cluster to 8 to fill the chip
        monkeypatch.setenv("SCMC_CLUSTER_FILL", "0")
        small = mk(); assert small.info()["resident_psi"] and small.info()["cluster"] == 4; small.close()
        monkeypatch.delenv("SCMC_CLUSTER_FILL")
        assert info["resident_psi"] and info["cluster"] == 8
    families = [(1, 2, info["resident"])] + ([(3, 4, info["resident_psi"])] if n == 1 else [])
    for m_stream, m_res, eligible in families:
        a = mk(); a.set_sweep_mode(m_stream)
        accs = [a.sweep(k, beta, sigma, hits=2)[0] for k in (1, 3)]
        finals[m_stream] = (a.get_state(), a.energies())
        if eligible:
            b = mk(); b.set_sweep_mode(m_res)
            for k, acc_a in zip((1, 3), accs):
                assert np.array_equal(acc_a, b.sweep(k, beta, sigma, hits=2)[0])
            assert np.array_equal(a.get_state(), b.get_state())
            assert np.array_equal(a.get_T(R - 1), b.get_T(R - 1))
            assert np.array_equal(a.energies(), b.energies())
            b.close()
        a.close()
    if 3 in finals and precision == 64:       # FP64: the families differ only at the 1e-16 level over 4 sweeps
        assert np.allclose(finals[1][0], finals[3][0], atol=1e-9) and np.allclose(finals[1][1], finals[3][1], rtol=1e-10)


def test_checkpoint_resume_continues_the_chain(scmc, tmp_path):
    """IO.jl checkpoint/resume: (state, seed, sweep counter) restore a run exactly -- the resumed chain is bit-identical to the
    uninterrupted one (the reference reseeds on resume; the counter-based stream fixes that)."""
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 6, 1, n_replicas=3, precision=32, seed=99)
    beta, sigma = np.array([1.0, 5.0, 25.0]), np.array([0.5, 0.2, 0.1])
    a = mk(); a.sweep(7, beta, sigma)
    b = mk(); b.sweep(3, beta, sigma)
    fn = str(tmp_path / "ckpt")
    scmc.checkpoint_(fn, b, None, current_sweep=3, βs=[1.0], energy=[0.0], σ=sigma, means={"energy": (b.energies() / 36, np.zeros(3))})
    c = scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 6, 1, n_replicas=3, precision=32, seed=12345)
    data = scmc.readCheckpoint(fn, c)
    assert data["state"].shape == (3, 4, 36) and int(data["current_sweep"]) == 3 and c.sweep_counter == 3
    c.sweep(4, beta, sigma)
    assert np.array_equal(a.get_state(), c.get_state())
    m, e = scmc.getmeans([fn, str(tmp_path / "missing")], "energy")
    assert m.shape == (1, 3)


def test_tensor_core_matvec_experiment_differs_by_rounding_only():
    """SCMC_MMA=1 (opt-in experiment): the 16x16 mat-vec of the streaming psi-gather kernel as 3xTF32 mma.sync over the
    32 sites of a warp.  Same stream, so after a few sweeps the acceptance counts and energies agree with the default FFMA path to FP32
    rounding; L = 10 has colour classes that are not multiples of 32 (tail lanes of the warp-uniform loop) and seam classes."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); import scmc_b200 as scmc\n"
        "out = []\n"
        "for L, R in ((10, 7), (24, 40)):\n"
        "    J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)\n"
        "    cfg = scmc.initializeCfg('nearest-neighbor', 'triangular', [J], L, 1, n_replicas=R, precision=32, seed=77)\n"
        "    cfg.set_sweep_mode(3)\n"
        "    acc, att = cfg.sweep(5, np.linspace(0.5, 8, R), np.full(R, 0.3))\n"
        "    out.append(dict(acc=acc.tolist(), att=att.tolist(), E=(cfg.energies() / cfg.nSites).tolist()))\n"
        "print(json.dumps(out))\n" % root)
    res = {}
    for m in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SCMC_MMA=m), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[m] = json.loads(r.stdout.strip().splitlines()[-1])
    for a, b in zip(res["0"], res["1"]):
        assert a["att"] == b["att"]
        acc0, acc1 = np.array(a["acc"], dtype=float), np.array(b["acc"], dtype=float)
        assert np.all(np.abs(acc0 - acc1) <= 2 + 1e-3 * acc0)              # a decision flips only on a near-tie
        assert np.allclose(a["E"], b["E"], rtol=0, atol=2e-3)             # ... and a flipped decision moves E/N by O(1/N)


def test_launch_writes_the_reference_result_files(scmc, tmp_path):
    """launch! ends in checkpoint! + saveMeans! (MonteCarlo.jl:103-104): one HDF5 file per temperature with `means/<key>/{mean,error}`,
    `rs`, `energy`, `σ`, `current_sweep`, `state` (d x N); getmeans / getenergies (IO.jl:64-143) collect them over the grid."""
    Ts = [0.1, 0.4, 1.6]
    fn = str(tmp_path / "run.h5")
    res = scmc.launch_(fn, "nearest-neighbor", 1, None, "triangular", 6, [NOTEBOOK_J], Ts, 2.0, 40, 60, measurement_rate=10, seed=3)
    assert res["files"] == [str(tmp_path / f"run_T{T:.6g}.h5") for T in Ts]
    psi = res["cfg"].get_state()
    for r, f in enumerate(res["files"]):
        one = scmc.readCheckpointH5(f)
        assert np.array_equal(one["state"], psi[r].T) and one["rs"].shape == (2, 36) and int(one["current_sweep"]) == 100
        assert one["energy"].shape == (4 + 6,) and float(one["σ"]) == res["sigma"][r]
        for key, (mean, err) in res["means"][r].items():
            assert np.array_equal(one["means"][key]["mean"], mean) and np.array_equal(one["means"][key]["error"], err)
    again = scmc.launch_(fn, "nearest-neighbor", 1, None, "triangular", 6, [NOTEBOOK_J], Ts, 2.0, 40, 60, measurement_rate=10, seed=3, overwrite=False)
    assert again["resumed"] and again["files"] == res["files"] and again["means"][2]["energy"][0] == res["means"][2]["energy"][0]
    name = lambda T: str(tmp_path / f"run_T{T:.6g}.h5")
    out = scmc.getmeansH5(name, str(tmp_path / "means.h5"), Ts)
    assert np.array_equal(out["energy/mean"], [m["energy"][0] for m in res["means"]]) and out["magnetizationVec/mean"].shape == (16, 3)
    en = scmc.getenergiesH5(name, str(tmp_path / "energies.h5"), Ts)
    assert np.array_equal(en["0.4"], scmc.readCheckpointH5(res["files"][1])["energy"])
    ann = scmc.launchAnnealing_(str(tmp_path / "sa.h5"), "nearest-neighbor", 1, "triangular", 6, [NOTEBOOK_J], 3.0, N_max=300, N_per_T=20,
                                min_acc_per_site=5, seed=5, n_restarts=2)
    assert len(ann["files"]) == 2
    assert np.allclose(scmc.readCheckpointH5(ann["files"][1])["state"], ann["state"][1].T, atol=0, rtol=0)


@pytest.mark.parametrize("precision", [32, 64])
def test_device_side_annealing_equals_host_driven_loop(scmc, precision):
    """runAnnealing!'s cooling loop inside the persistent resident kernel (one CTA per replica) vs the same loop driven from the
    host one sweep at a time: identical final beta, sigma, sweep counts and bit-identical states for every replica."""
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 12, 1, n_replicas=24, precision=precision, seed=4321)
    kw = dict(T_fac=0.9, N_max=700, N_o=0, N_per_T=20, min_acc_per_site=3, min_accrate=0.02, σ_min=0.05, log_rate=0)      # no E / beta series: launches are not cut at log points
    a = mk(); a.set_sweep_mode(3)                  # streaming psi-gather kernel, host-driven controller
    ra = scmc.runAnnealing_(a, 2.0, **kw)
    b = mk()                                       # auto: psi family, single-CTA resident kernel, device-side controller
    assert b.info()["resident_psi"] and b.info()["cluster"] == 1
    l0 = b.info()["launches"]
    rb = scmc.runAnnealing_(b, 2.0, **kw)
    assert b.info()["launches"] - l0 <= 12 and a.info()["launches"] > 500      # the controller really ran inside the persistent kernel
    assert np.array_equal(ra["sweeps"], rb["sweeps"]) and ra["sweeps"].min() > 50
    assert np.array_equal(ra["beta"], rb["beta"]) and np.array_equal(ra["sigma"], rb["sigma"])
    assert np.array_equal(a.get_state(), b.get_state())
    assert np.array_equal(ra["E"], rb["E"])
    assert len(set(ra["sweeps"].tolist())) > 1 or ra["sweeps"][0] == 699      # replicas stop independently (or all hit N_max)


@pytest.mark.parametrize("mode,stream", [(0, 1), (1, 1), (2, 2), (3, 2), (4, 1), (0, 2)])
@pytest.mark.parametrize("lattice,L", [("square", 5), ("cubic", 3), ("chain", 7), ("square", 4)])
def test_other_lattices_follow_the_oracle_chain(scmc, oracle_mod, lattice, L, mode, stream):
    """Neighbour counts that are not a multiple of 3 (padding slots -> zero site), odd sizes (greedy colouring), 1-D and 3-D
    lattices: energies and the FP64 counter-based chain of every sweep kernel against the oracle."""
    seed = 31415
    cfg = scmc.initializeCfg("nearest-neighbor", lattice, [random_symmetric_J(5)], L, 1, n_replicas=2, precision=64, seed=seed, replica_offset=3, stream=stream)
    info = cfg.info()
    if (mode == 2 and not info["resident"]) or (mode == 4 and not info["resident_psi"]):
        pytest.skip("resident kernel of this family not available for the model")
    cfg.set_sweep_mode(mode)
    o = make_oracle(cfg, oracle_mod)
    o.randomize_v1(seed, 4)
    assert np.allclose(cfg.get_state(1, 1)[0], o.get_state(), atol=1e-13)
    assert abs(cfg.energies()[1] - o.energy()) < 1e-11 * cfg.nSites
    acc, _ = cfg.sweep(4, np.array([1.0, 3.0]), np.array([0.5, 0.3]))
    a = o.sweeps_colour(stream, 4, 2, cfg.colour, 3.0, 0.3, seed, 4, 0)
    assert acc[1] == a
    assert np.allclose(cfg.get_state(1, 1)[0], o.get_state(), atol=1e-10)
    assert abs(cfg.energies()[1] - o.energy()) < 1e-9 * cfg.nSites


def test_production_fp32_psi_family_matches_fp64_reference_arithmetic(scmc):
    """Cross-family, cross-precision statistics: the production path (FP32, psi-gather kernels, SFU transcendentals, 24-bit
    uniforms) against the reference-shaped arithmetic (FP64, T-gather kernels) on triangular 12x12 at three temperatures.
    Independent chains (different seeds): energies, specific heats and |m_sigma| agree within their error bars."""
    Ts = np.repeat([0.05, 0.3, 1.2], 24)
    out = {}
    for tag, precision, mode, seed in (("fp32_psi", 32, 0, 101), ("fp64_T", 64, 1, 202)):
        cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 12, 1, n_replicas=len(Ts), precision=precision, seed=seed)
        cfg.set_sweep_mode(mode)
        res = scmc.run_(cfg, None, Ts, 2.0, 3000, 4000, measurement_rate=10)
        m = res["mean"].reshape(3, 24, -1)
        e = m[:, :, 0]
        c = (m[:, :, 1] - e ** 2) * cfg.nSites / np.array([0.05, 0.3, 1.2])[:, None] ** 2
        out[tag] = dict(e=(e.mean(1), e.std(1, ddof=1) / np.sqrt(24)), c=(c.mean(1), c.std(1, ddof=1) / np.sqrt(24)),
                        ms=(m[:, :, 2].mean(1), m[:, :, 2].std(1, ddof=1) / np.sqrt(24)), acc=res["acc_rate"].reshape(3, 24).mean(1))
    for key in ("e", "c", "ms"):
        a, b = out["fp32_psi"][key], out["fp64_T"][key]
        z = np.abs(a[0] - b[0]) / np.hypot(a[1], b[1])
        assert np.all(z < 5.0), (key, a, b, z)
    assert np.all(np.abs(out["fp32_psi"]["acc"] - out["fp64_T"]["acc"]) < 0.05)
    assert abs(out["fp32_psi"]["e"][0][0] - (-3.6 + 3 * 0.05)) < 0.01          # low-T equipartition anchor


@pytest.mark.parametrize("precision", [32, 64])
def test_device_side_run_equals_host_driven_schedule(scmc, precision):
    """run! inside the persistent resident kernel (beta ramp, sigma adaptation, measure! rows) vs the host-driven loop: identical
    chains (states, adapted sigma, acceptance), measurement rows equal to rounding (E re-summed from the resident states)."""
    R = 10
    Ts = np.geomspace(0.05, 1.5, R)
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 12, 1, n_replicas=R, precision=precision, seed=2468)
    a = mk(); a.set_sweep_mode(3)              # streaming psi-gather kernel, host-driven schedule
    ra = scmc.run_(a, None, Ts, 2.0, 230, 345, measurement_rate=10)
    b = mk()                                   # auto: single-CTA resident psi kernel, device-side schedule
    assert b.info()["resident_psi"] and b.info()["cluster"] == 1
    l0 = b.info()["launches"]
    rb = scmc.run_(b, None, Ts, 2.0, 230, 345, measurement_rate=10)
    assert b.info()["launches"] - l0 <= 6      # 575 sweeps in a handful of persistent launches: the schedule really ran on the device
    assert a.info()["launches"] > 100
    assert ra["rows"] == rb["rows"] == 34
    assert ra["therm_energy"].shape == rb["therm_energy"].shape == (23, R)
    assert np.array_equal(ra["sigma"], rb["sigma"]) and np.array_equal(ra["acc_rate"], rb["acc_rate"])
    assert np.array_equal(a.get_state(), b.get_state())
    tol = 1e-10 if precision == 64 else 2e-5
    assert np.allclose(ra["series"], rb["series"], rtol=tol, atol=tol)
    assert np.allclose(ra["mean"], rb["mean"], rtol=tol, atol=tol)
    assert np.allclose(ra["therm_energy"], rb["therm_energy"], rtol=tol, atol=tol)     # E/N logged at the sigma-adaptation points
    # a resumed run (current_sweep) continues both paths identically
    ra2 = scmc.run_(a, None, Ts, 2.0, 230, 500, measurement_rate=10, current_sweep=575, σ=ra["sigma"])
    rb2 = scmc.run_(b, None, Ts, 2.0, 230, 500, measurement_rate=10, current_sweep=575, σ=rb["sigma"])
    assert ra2["rows"] == rb2["rows"] and np.array_equal(a.get_state(), b.get_state())


def test_more_replicas_than_the_grid_y_limit(scmc):
    """70 000 replicas of a tiny lattice: launches are chunked over the 65 535 gridDim.y limit, and a replica beyond the boundary
    follows exactly the chain of a single-replica handle with the same global id."""
    R = 70000
    big = scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 3, 1, n_replicas=R, precision=32, seed=7)
    big.set_sweep_mode(3)
    big.set_lanes(1)            # one stream: all 70 000 replicas in one chunked launch sequence
    acc, att = big.sweep(3, 2.0, 0.4)
    assert np.all(att == 3 * 2 * 9) and acc.min() > 0
    for rid in (65534, 65535, 69999):
        one = scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 3, 1, n_replicas=1, precision=32, seed=7, replica_offset=rid)
        one.set_sweep_mode(3)
        a1, _ = one.sweep(3, 2.0, 0.4)
        assert a1[0] == acc[rid]
        assert np.array_equal(one.get_state()[0], big.get_state(rid, 1)[0])
    assert np.isfinite(big.energies()).all() and big.measure().shape == (R, 21)


@pytest.mark.gpu
@pytest.mark.parametrize("mode, n, labels", [(3, 1, 1), (1, 1, 1), (1, 2, 3)])
def test_replica_lanes_do_not_change_chains(scmc, mode, n, labels):
    """Two replica lanes on side streams (scmc_set_lanes) against everything on the handle's stream: same accepted counts, same
    states, for both streaming kernel families and an odd replica count; work queued on the handle's stream afterwards sees the
    joined result."""
    def mk():
        if labels == 1:
            c = scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 12, n, n_replicas=37, precision=32, seed=11)
        else:
            c = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 12, n, n_replicas=37, precision=32, seed=11)
        c.set_sweep_mode(mode)
        return c
    a, b = mk(), mk()
    a.set_lanes(1); b.set_lanes(2)
    beta = np.linspace(0.5, 20.0, 37)
    for _ in range(3):
        acc_a, _ = a.sweep(5, beta, 0.5)
        acc_b, _ = b.sweep(5, beta, 0.5)
        assert np.array_equal(acc_a, acc_b)
        assert np.array_equal(a.energies(), b.energies())
    assert np.array_equal(a.get_state(), b.get_state())
    assert b.info()["launches"] > a.info()["launches"]


@pytest.mark.gpu
@pytest.mark.parametrize("n, precision", [(2, 32), (1, 32), (2, 64)])
def test_label_major_field_is_bit_identical(scmc, monkeypatch, n, precision):
    """Several labels with the same number of bonds per label at every site (TG/h-BN): the label-major gather (one label at a time, sum in
    registers) against the per-thread-label gather of the same build (SCMC_LABEL_MAJOR=0): same loads, additions and FMAs in the same
    order, so acceptance counts, local fields and states must agree bit for bit.  Energies come from k_measure, whose mat-vec differs between
    the two paths (couplings from shared memory with packed FMAs vs from the constant bank), so they agree to rounding only."""
    def mk(flag):
        monkeypatch.setenv("SCMC_LABEL_MAJOR", flag)
        c = scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 12, n, n_replicas=16, precision=precision, seed=99)
        c.set_sweep_mode(1)
        return c
    a, b = mk("0"), mk("1")
    assert np.array_equal(a.get_state(), b.get_state())
    etol = 1e-12 if precision == 64 else 2e-6
    assert np.allclose(a.energies(), b.energies(), rtol=etol, atol=0)
    assert np.array_equal(a.local_field(3), b.local_field(3))
    beta = np.linspace(0.5, 10.0, 16)
    for _ in range(3):
        acc_a, _ = a.sweep(4, beta, 0.4)
        acc_b, _ = b.sweep(4, beta, 0.4)
        assert np.array_equal(acc_a, acc_b)
    assert np.array_equal(a.get_state(), b.get_state())
    assert np.allclose(a.energies(), b.energies(), rtol=etol, atol=0)
    assert np.allclose(a.measure(), b.measure(), rtol=10 * etol, atol=10 * etol)


@pytest.mark.gpu
@pytest.mark.parametrize("L, cmin, precision", [(12, 2, 32), (12, 4, 32), (12, 2, 64), (96, 1, 32)])
def test_device_side_run_in_clusters_equals_host_driven_schedule(scmc, monkeypatch, L, cmin, precision):
    """run! inside the persistent kernel when a replica spans a cluster of CTAs (acceptance counts and measurement sums reduced through
    distributed shared memory) vs the host-driven schedule on the streaming kernel: identical chains, adapted sigma and acceptance rates;
    measurement rows and the thermalisation energy log equal to rounding.  L = 12 forces a larger cluster than the lattice needs
    (SCMC_CLUSTER_MIN), L = 96 needs a cluster of 2 by size."""
    R = 6 if L == 12 else 3
    n_th, n_me = (120, 150) if L == 12 else (30, 40)
    Ts = np.geomspace(0.05, 1.5, R)
    def mk(c):
        monkeypatch.setenv("SCMC_CLUSTER_MIN", str(c))
        return scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], L, 1, n_replicas=R, precision=precision, seed=1357)
    a = mk(1); a.set_sweep_mode(3)              # streaming psi-gather kernel, host-driven schedule
    ra = scmc.run_(a, None, Ts, 2.0, n_th, n_me, measurement_rate=10)
    b = mk(cmin)
    assert b.info()["resident_psi"] and b.info()["cluster"] >= max(2, cmin)
    l0 = b.info()["launches"]
    rb = scmc.run_(b, None, Ts, 2.0, n_th, n_me, measurement_rate=10)
    assert b.info()["launches"] - l0 <= 6      # the schedule ran inside the persistent cluster kernel
    assert ra["rows"] == rb["rows"] == (n_th + n_me) // 10 - n_th // 10
    assert np.array_equal(ra["sigma"], rb["sigma"]) and np.array_equal(ra["acc_rate"], rb["acc_rate"])
    assert np.array_equal(a.get_state(), b.get_state())
    tol = 1e-10 if precision == 64 else 2e-5
    assert np.allclose(ra["series"], rb["series"], rtol=tol, atol=tol)
    assert np.allclose(ra["mean"], rb["mean"], rtol=tol, atol=tol)
    assert np.allclose(ra["therm_energy"], rb["therm_energy"], rtol=tol, atol=tol)
    # resumed run: the window's accepted count is carried over between launches by rank 0 only
    ra2 = scmc.run_(a, None, Ts, 2.0, n_th, n_me + 55, measurement_rate=10, current_sweep=n_th + n_me, σ=ra["sigma"])
    rb2 = scmc.run_(b, None, Ts, 2.0, n_th, n_me + 55, measurement_rate=10, current_sweep=n_th + n_me, σ=rb["sigma"])
    assert ra2["rows"] == rb2["rows"] and np.array_equal(a.get_state(), b.get_state())
    assert np.array_equal(ra2["acc_rate"], rb2["acc_rate"])


@pytest.mark.gpu
@pytest.mark.parametrize("cmin, precision", [(2, 32), (4, 32), (2, 64)])
def test_device_side_annealing_in_clusters_equals_host_driven_loop(scmc, monkeypatch, cmin, precision):
    """runAnnealing!'s cooling loop inside the persistent kernel when a replica spans a cluster of CTAs (the sweep's accepted count is
    reduced through distributed shared memory every sweep) vs the loop driven from the host one sweep at a time: identical sweep counts,
    final beta and sigma, and bit-identical states for every replica."""
    def mk(c):
        monkeypatch.setenv("SCMC_CLUSTER_MIN", str(c))
        return scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 12, 1, n_replicas=10, precision=precision, seed=4321)
    kw = dict(T_fac=0.9, N_max=700, N_o=0, N_per_T=20, min_acc_per_site=3, min_accrate=0.02, σ_min=0.05, log_rate=0)      # no E / beta series: launches are not cut at log points
    a = mk(1); a.set_sweep_mode(3)                 # streaming psi-gather kernel, host-driven controller
    ra = scmc.runAnnealing_(a, 2.0, **kw)
    b = mk(cmin)
    assert b.info()["resident_psi"] and b.info()["cluster"] == cmin
    l0 = b.info()["launches"]
    rb = scmc.runAnnealing_(b, 2.0, **kw)
    assert b.info()["launches"] - l0 <= 12         # the controller ran inside the persistent cluster kernel
    assert np.array_equal(ra["sweeps"], rb["sweeps"]) and ra["sweeps"].min() > 50
    assert np.array_equal(ra["beta"], rb["beta"]) and np.array_equal(ra["sigma"], rb["sigma"])
    assert np.array_equal(a.get_state(), b.get_state())
    etol = 1e-12 if precision == 64 else 2e-6
    assert np.allclose(ra["E"], rb["E"], rtol=etol, atol=0)


def _t_family_model(scmc, kind, R, precision, seed):
    if kind == "tghbn_n2":
        return scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 12, 2, n_replicas=R, precision=precision, seed=seed)
    if kind == "tghbn_n1":
        return scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.1, 0.5], 12, 1, n_replicas=R, precision=precision, seed=seed)
    if kind == "honey_n2":          # d = 6 without on-site term: T-gather family in auto mode
        return scmc.initializeCfg("nearest-neighbor", "honeycomb", [np.eye(16)], 8, 2, n_replicas=R, precision=precision, seed=seed)
    if kind == "tri_forced":        # psi-family model driven through the T-gather kernels (modes 1 / 2)
        return scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 12, 1, n_replicas=R, precision=precision, seed=seed)
    raise KeyError(kind)


@pytest.mark.gpu
@pytest.mark.parametrize("kind, cmin, precision", [("tghbn_n2", 1, 32), ("tghbn_n2", 2, 32), ("tghbn_n1", 1, 32), ("honey_n2", 1, 32), ("honey_n2", 2, 64),
                                                   ("tri_forced", 1, 32), ("tghbn_n2", 1, 64)])
def test_t_family_device_side_run_equals_host_driven_schedule(scmc, monkeypatch, kind, cmin, precision):
    """run! inside the persistent T-gather kernel (one CTA or a cluster per replica) vs the host-driven schedule on the streaming T-gather
    kernel: identical chains, adapted sigma and acceptance rates; measurement rows and the thermalisation log equal to rounding."""
    R = 6
    Ts = np.geomspace(0.05, 1.5, R)
    def mk(c):
        monkeypatch.setenv("SCMC_CLUSTER_MIN", str(c))
        return _t_family_model(scmc, kind, R, precision, 2024)
    a = mk(1); a.set_sweep_mode(1)              # streaming T-gather kernel, host-driven schedule
    ra = scmc.run_(a, None, Ts, 2.0, 120, 150, measurement_rate=10)
    b = mk(cmin)
    if kind == "tri_forced": b.set_sweep_mode(2)
    assert b.info()["resident"] and b.info()["cluster"] == cmin
    l0 = b.info()["launches"]
    rb = scmc.run_(b, None, Ts, 2.0, 120, 150, measurement_rate=10)
    assert b.info()["launches"] - l0 <= 8      # the schedule ran inside the persistent kernel
    assert a.info()["launches"] > 100
    assert ra["rows"] == rb["rows"] == 15
    assert np.array_equal(ra["sigma"], rb["sigma"]) and np.array_equal(ra["acc_rate"], rb["acc_rate"])
    assert np.array_equal(a.get_state(), b.get_state())
    assert np.array_equal(a.get_T(R - 1), b.get_T(R - 1))
    tol = 1e-10 if precision == 64 else 2e-5
    assert np.allclose(ra["series"], rb["series"], rtol=tol, atol=tol)
    assert np.allclose(ra["mean"], rb["mean"], rtol=tol, atol=tol)
    assert np.allclose(ra["therm_energy"], rb["therm_energy"], rtol=tol, atol=tol)
    ra2 = scmc.run_(a, None, Ts, 2.0, 120, 205, measurement_rate=10, current_sweep=270, σ=ra["sigma"])
    rb2 = scmc.run_(b, None, Ts, 2.0, 120, 205, measurement_rate=10, current_sweep=270, σ=rb["sigma"])
    assert ra2["rows"] == rb2["rows"] and np.array_equal(a.get_state(), b.get_state())
    assert np.array_equal(ra2["acc_rate"], rb2["acc_rate"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind, cmin, precision", [("tghbn_n2", 1, 32), ("tghbn_n2", 2, 32), ("honey_n2", 1, 64), ("tri_forced", 2, 32)])
def test_t_family_device_side_annealing_equals_host_driven_loop(scmc, monkeypatch, kind, cmin, precision):
    """runAnnealing!'s cooling loop inside the persistent T-gather kernel vs the loop driven from the host one sweep at a time."""
    def mk(c):
        monkeypatch.setenv("SCMC_CLUSTER_MIN", str(c))
        return _t_family_model(scmc, kind, 8, precision, 777)
    kw = dict(T_fac=0.9, N_max=500, N_o=0, N_per_T=20, min_acc_per_site=3, min_accrate=0.02, σ_min=0.05, log_rate=0)
    a = mk(1); a.set_sweep_mode(1)
    ra = scmc.runAnnealing_(a, 2.0, **kw)
    b = mk(cmin)
    if kind == "tri_forced": b.set_sweep_mode(2)
    assert b.info()["resident"] and b.info()["cluster"] == cmin
    l0 = b.info()["launches"]
    rb = scmc.runAnnealing_(b, 2.0, **kw)
    assert b.info()["launches"] - l0 <= 12 and a.info()["launches"] > 300
    assert np.array_equal(ra["sweeps"], rb["sweeps"]) and ra["sweeps"].min() > 50
    assert np.array_equal(ra["beta"], rb["beta"]) and np.array_equal(ra["sigma"], rb["sigma"])
    assert np.array_equal(a.get_state(), b.get_state())
    etol = 1e-12 if precision == 64 else 2e-6
    assert np.allclose(ra["E"], rb["E"], rtol=etol, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["psi_tri48", "T_honey24"])
def test_chip_filling_cluster_rule_does_not_change_chains(scmc, monkeypatch, kind):
    """Few replicas: scmc_create grows the cluster beyond what the lattice needs so that idle SMs get work (SCMC_CLUSTER_FILL=0 switches the
    rule off).  The cluster size is a layout choice: run!, annealing and plain sweeps must give identical chains with the rule on and off."""
    def mk(fill, R):
        monkeypatch.setenv("SCMC_CLUSTER_FILL", fill)
        if kind == "psi_tri48":
            return scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 48, 1, n_replicas=R, precision=32, seed=31)
        return scmc.initializeCfg("nearest-neighbor", "honeycomb", [np.eye(16)], 24, 2, n_replicas=R, precision=32, seed=31)
    R = 12
    a, b = mk("0", R), mk("1", R)
    assert a.info()["cluster"] == 1 and b.info()["cluster"] >= 2          # the rule triggers for this shape
    Ts = np.geomspace(0.05, 1.5, R)
    ra = scmc.run_(a, None, Ts, 2.0, 60, 60, measurement_rate=10)
    rb = scmc.run_(b, None, Ts, 2.0, 60, 60, measurement_rate=10)
    assert np.array_equal(ra["sigma"], rb["sigma"]) and np.array_equal(ra["acc_rate"], rb["acc_rate"])
    assert np.array_equal(a.get_state(), b.get_state())
    assert np.allclose(ra["mean"], rb["mean"], rtol=2e-5, atol=2e-5)
    acc_a, _ = a.sweep(7, 1.0 / Ts, ra["sigma"]); acc_b, _ = b.sweep(7, 1.0 / Ts, rb["sigma"])
    assert np.array_equal(acc_a, acc_b) and np.array_equal(a.get_state(), b.get_state())
    a2, b2 = mk("0", R), mk("1", R)
    kw = dict(T_fac=0.9, N_max=300, N_o=0, N_per_T=20, min_acc_per_site=3, min_accrate=0.02, σ_min=0.05)
    qa, qb = scmc.runAnnealing_(a2, 2.0, **kw), scmc.runAnnealing_(b2, 2.0, **kw)
    assert np.array_equal(qa["sweeps"], qb["sweeps"]) and np.array_equal(qa["beta"], qb["beta"]) and np.array_equal(qa["sigma"], qb["sigma"])
    assert np.array_equal(a2.get_state(), b2.get_state())


@pytest.mark.gpu
def test_oversized_scratch_buffers_of_run_are_independent(scmc):
    """Thermalisation log and measurement rows both above the 64 MB threshold of the scratch pool (4096 replicas of a 3x3 lattice, a point
    every sweep): the log buffer goes out of scope before the measurement phase and must release only its own slot.  Device-side schedule
    against the host-driven one: same chains, rows and log equal to rounding."""
    R = 4096
    Ts = np.geomspace(0.1, 2.0, R)
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 3, 1, n_replicas=R, precision=32, seed=8)
    a = mk(); a.set_sweep_mode(3)
    ra = scmc.run_(a, None, Ts, 2.0, 2100, 100, measurement_rate=1)
    b = mk()
    rb = scmc.run_(b, None, Ts, 2.0, 2100, 100, measurement_rate=1)
    assert ra["rows"] == rb["rows"] == 100 and ra["therm_energy"].shape == rb["therm_energy"].shape == (2100, R)
    assert ra["therm_energy"].nbytes > (64 << 20) and ra["series"].nbytes > (64 << 20)
    assert np.array_equal(a.get_state(), b.get_state()) and np.array_equal(ra["sigma"], rb["sigma"])
    assert np.isfinite(rb["series"]).all() and np.isfinite(rb["therm_energy"]).all()
    assert np.allclose(ra["series"], rb["series"], rtol=2e-5, atol=2e-5)
    assert np.allclose(ra["therm_energy"], rb["therm_energy"], rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("lattice, L, n, precision", [("triangular", 12, 1, 32), ("honeycomb", 8, 2, 32), ("honeycomb", 6, 1, 64), ("square", 7, 1, 64)])
def test_structure_factor_grid_equals_direct_sum(scmc, lattice, L, n, precision):
    """S^a(k) on the allowed momenta of the periodic cluster through the separable DFT over the cell indices (scmc_structure_factor_grid)
    against the direct sum over sites per momentum (scmc_structure_factor): same FP64 accumulation of the same T values.  The box reaches
    beyond the first reciprocal cell, where the basis phases of a two-site cell matter; non-allowed momenta fall back to the direct sum."""
    R = 3
    J = [np.eye(16)] if n == 2 else [random_symmetric_J(4)]
    cfg = scmc.initializeCfg("nearest-neighbor", lattice, J, L, n, n_replicas=R, precision=precision, seed=17)
    cfg.sweep(3, np.array([0.5, 2.0, 8.0]), np.full(R, 0.4))
    ks = scmc.getKsInBox(lattice, L, [5 * np.pi, 5 * np.pi], [0.3, -0.2])
    assert len(ks) > L * L                                   # more momenta than one reciprocal cell holds
    direct = scmc.structureFactor(cfg, ks, method="direct")
    grid = scmc.structureFactor(cfg, ks, method="grid")
    auto = scmc.structureFactor(cfg, ks)
    scale = np.abs(direct).max()
    assert scale > 1e-3 and np.allclose(grid, direct, rtol=1e-9, atol=1e-10 * scale)
    assert np.array_equal(auto, grid)
    # component sums on the device (computeStructureFactor with `components`; default 0..14 = the reference's 1:15)
    ssum = scmc.structureFactorSum(cfg, ks)
    assert ssum.shape == (R, len(ks)) and np.allclose(ssum, direct[:, :15, :].sum(axis=1), rtol=1e-9, atol=1e-10 * scale)
    assert np.allclose(scmc.structureFactorSum(cfg, ks, components=[4, 8, 12]), direct[:, [4, 8, 12], :].sum(axis=1), rtol=1e-9, atol=1e-10 * scale)
    off = ks[:5] + 1e-3                                       # not allowed momenta
    assert np.array_equal(scmc.structureFactor(cfg, off), scmc.structureFactor(cfg, off, method="direct"))
    assert np.allclose(scmc.structureFactorSum(cfg, off), scmc.structureFactor(cfg, off, method="direct")[:, :15, :].sum(axis=1), rtol=1e-12)
    with pytest.raises(ValueError):
        scmc.structureFactor(cfg, off, method="grid")


@pytest.mark.gpu
@pytest.mark.parametrize("lattice, L, n", [("honeycomb", 8, 2), ("triangular", 12, 1)])
def test_structure_factor_running_sum_on_the_device(scmc, lattice, L, n):
    """StructureFactorPlan: S(k) summed over components, accumulated on the device over measurements and fetched once, against the host-side
    average of structureFactorSum taken at the same points of the chain."""
    R = 4
    J = [np.eye(16)] if n == 2 else [NOTEBOOK_J]
    cfg = scmc.initializeCfg("nearest-neighbor", lattice, J, L, n, n_replicas=R, precision=32, seed=23)
    ks = scmc.getKsInBox(lattice, L, [4 * np.pi, 4 * np.pi], [0.0, 0.0])
    plan = scmc.StructureFactorPlan(cfg, ks)
    assert plan.mean(reset=False)[1] == 0                      # nothing accumulated yet
    beta, sigma = np.array([0.5, 1.0, 3.0, 9.0]), np.full(R, 0.4)
    host = np.zeros((R, len(ks))); n_meas = 5
    for _ in range(n_meas):
        cfg.sweep(3, beta, sigma)
        plan.accumulate()
        host += scmc.structureFactorSum(cfg, ks)
    with pytest.raises(scmc.ScmcError):                        # another momentum list while a sum is running
        scmc.StructureFactorPlan(cfg, ks[:-3]).accumulate()
    mean, cnt = plan.mean(reset=True)
    assert cnt == n_meas and np.allclose(mean, host / n_meas, rtol=1e-12, atol=1e-12 * np.abs(host).max())
    assert plan.mean()[1] == 0                                 # reset dropped the running sum
    other = scmc.StructureFactorPlan(cfg, ks[:-3], components=[4, 8, 12])
    other.accumulate()
    m2, c2 = other.mean()
    assert c2 == 1 and np.allclose(m2, scmc.structureFactorSum(cfg, ks[:-3], components=[4, 8, 12]), rtol=1e-12)


@pytest.mark.parametrize("case", [
    ("triangular", 128, 37),     # BASELINE C5 lattice: 3 big classes (TMA tiles) + seam classes (gather); 37 replicas = ragged replica chunks
    ("triangular", 48, 19),      # L % 3 == 0: three classes of 768 sites, 6 tiles each
    ("honeycomb", 24, 21),       # z = 3
    ("square", 32, 18),          # z = 4, padded to 6 slots (zero-site neighbours)
])
@pytest.mark.parametrize("stream", [1, 2])
def test_tma_pipelined_kernel_is_bit_identical(scmc, case, stream):
    """The TMA-pipelined streaming kernel (mode 5: cp.async.bulk + mbarrier ring, scmc_tma.cu) realises the chain of the direct-gather
    kernel (mode 3) bit for bit: same stream, same summation order.  Reference work unit: localUpdate!, MonteCarlo.jl:263-286."""
    lattice, L, R = case
    mk = lambda: scmc.initializeCfg("nearest-neighbor", lattice, [NOTEBOOK_J], L, 1, n_replicas=R, precision=32, seed=4242, replica_offset=3, stream=stream)
    beta = np.geomspace(0.5, 50.0, R); sigma = np.geomspace(0.8, 0.03, R)
    a = mk(); a.set_sweep_mode(3)
    b = mk(); b.set_sweep_mode(5)
    assert b.info()["tma"], "the class tiles of this lattice must decompose into bulk copies"
    for lanes in (1, 2):
        a.set_lanes(lanes); b.set_lanes(lanes)
        for k in (1, 3):
            hits = 2 if k == 1 else 3                     # 3 hits per visit: a visit straddles two groups of the v2 acceptance uniforms
            acc_a, att_a = a.sweep(k, beta, sigma, hits=hits)
            acc_b, att_b = b.sweep(k, beta, sigma, hits=hits)
            assert np.array_equal(acc_a, acc_b) and np.array_equal(att_a, att_b)
            assert np.array_equal(a.get_state(), b.get_state())
    assert np.array_equal(a.energies(), b.energies())
    assert acc_a.min() > 0
    a.close(); b.close()


@pytest.mark.parametrize("lattice,L,n", [("triangular", 12, 1), ("triangular", 128, 1), ("square", 16, 1), ("honeycomb", 8, 1), ("triangular", 9, 3)])
@pytest.mark.parametrize("precision", [64, 32])
def test_forward_bond_measurement_equals_the_full_sum(scmc, oracle_mod, monkeypatch, lattice, L, n, precision):
    """One coupling label (J = J^T): scmc_create looks for z / 2 "forward" neighbour slots that hold every bond at exactly one of its ends (three of
    the six directions of the triangular lattice, two of four on the square lattice; none on the honeycomb lattice, whose sublattices differ), and
    the psi-family measurement then sums each bond once instead of twice with a factor 1/2 (getEnergy, src/Configuration.jl:236-253).  Same
    energies and magnetisation as the full sum (SCMC_MEASURE_FWD=0), and as the oracle on a small lattice."""
    R = 5
    cfg = scmc.initializeCfg("nearest-neighbor", lattice, [random_symmetric_J(21)], L, n, n_replicas=R, precision=precision, seed=31)
    tol = 1e-12 if precision == 64 else 2e-6
    rows = cfg.measure().copy(); e = cfg.energies().copy()
    monkeypatch.setenv("SCMC_MEASURE_FWD", "0")
    rows0 = cfg.measure().copy(); e0 = cfg.energies().copy()
    monkeypatch.delenv("SCMC_MEASURE_FWD")
    assert np.allclose(e, e0, rtol=tol, atol=tol * cfg.nSites) and np.allclose(rows, rows0, rtol=tol, atol=tol)
    assert np.array_equal(rows[:, 5:], rows0[:, 5:])                      # the magnetisation sums do not depend on the bond set
    if cfg.nSites <= 200:
        o = make_oracle(cfg, oracle_mod, replica=2)
        assert abs(o.energy() - e[2]) < (1e-10 if precision == 64 else 2e-5) * cfg.nSites
    cfg.close()


@pytest.mark.gpu
def test_compile_time_stage_capacity_is_the_same_chain(scmc, monkeypatch):
    """The production instantiation of the tensor-core TMA kernel has its stage capacity at compile time (640 slots: strides are immediates of the
    shared-memory loads); lattices whose tile plan needs more fall back to the run-time form.  Both must realise the same chain bit for bit
    (SCMC_TMA_FIXCAP=0 forces the run-time form where the compile-time one applies)."""
    R = 21
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 128, 1, n_replicas=R, precision=32, seed=99, replica_offset=2, stream=2)
    beta = np.geomspace(0.5, 50.0, R); sigma = np.geomspace(0.8, 0.03, R)
    a = mk(); a.set_sweep_mode(6)
    acc_a, _ = a.sweep(3, beta, sigma, hits=2)
    monkeypatch.setenv("SCMC_TMA_FIXCAP", "0")
    b = mk(); b.set_sweep_mode(6)
    acc_b, _ = b.sweep(3, beta, sigma, hits=2)
    monkeypatch.delenv("SCMC_TMA_FIXCAP")
    assert np.array_equal(acc_a, acc_b) and np.array_equal(a.get_state(), b.get_state()) and acc_a.min() > 0
    a.close(); b.close()


@pytest.mark.parametrize("case", [
    ("triangular", 128, 37, 1),  # BASELINE C5 lattice, ragged replica chunks
    ("triangular", 48, 19, 1),   # L % 3 == 0
    ("honeycomb", 24, 21, 1),    # z = 3
    ("square", 32, 18, 1),       # z = 4, padded to 6 slots
    ("triangular", 48, 9, 3),    # n = 3 (d = 4, second generator table of the density basis)
])
@pytest.mark.parametrize("stream", [1, 2])
@pytest.mark.parametrize("coupling", ["notebook", "dense"])
def test_tensor_core_variant_follows_the_ffma_chain(scmc, case, stream, coupling):
    """Mode 6 (what auto runs for streaming psi-gather passes in FP32): the TMA-pipelined kernel with the 16x16 mat-vec of a visit on the
    tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, FP32 accumulation in TMEM) and packed FFMA2 density sums.  Its rounding differs from the
    FFMA order of mode 5, so the chains agree to rounding: from identical states, after whole sweeps almost every site holds the same
    state (a near-tie in an accept test may flip a handful), accept counts agree to a few, energies to 1e-5.  localUpdate!, MonteCarlo.jl:263-286."""
    lattice, L, R, n = case
    J = NOTEBOOK_J if coupling == "notebook" else random_symmetric_J(11)
    mk = lambda: scmc.initializeCfg("nearest-neighbor", lattice, [J], L, n, n_replicas=R, precision=32, seed=777, replica_offset=5, stream=stream)
    beta = np.geomspace(0.5, 50.0, R); sigma = np.geomspace(0.8, 0.03, R)
    a = mk(); a.set_sweep_mode(5)
    b = mk(); b.set_sweep_mode(6)
    assert b.info()["tma"] and b.info()["tma_tc"] and not a.info()["tma_tc"]
    a.set_sweep_mode(3); assert not a.info()["tma_tc"]
    a.set_sweep_mode(5)
    for hits in (2, 3):                                   # 2: unrolled hits, random numbers generated under the MMA; 3: run-time hit loop
        acc_a, att_a = a.sweep(1, beta, sigma, hits=hits)
        acc_b, att_b = b.sweep(1, beta, sigma, hits=hits)
        assert np.array_equal(att_a, att_b)
        assert np.abs(acc_a.astype(np.int64) - acc_b.astype(np.int64)).max() <= 3
        sa, sb = a.get_state(), b.get_state()
        differing = np.abs(sa - sb).reshape(R, a.nSites, -1).max(axis=2) > 1e-4
        assert differing.mean() < 1e-4, differing.sum()
        Ea, Eb = a.energies(), b.energies()
        same = ~differing.any(axis=1)                     # replicas without a flipped near-tie: the energies agree to rounding
        assert same.sum() >= R - 3 and np.allclose(Ea[same], Eb[same], rtol=2e-5, atol=2e-6 * a.nSites)
        b.set_state(sa)                                   # continue both from the same states: differences do not accumulate in the test
    a.close(); b.close()


@pytest.mark.parametrize("mode", [5, 6])
@pytest.mark.parametrize("stream", [1, 2])
def test_fp32_tma_kernels_follow_the_oracle_chain(scmc, oracle_mod, mode, stream):
    """The FP32 production kernels of the headline workload (mode 5: TMA-pipelined, mode 6: + tensor-core mat-vec) against the oracle's FP64
    restatement of the same colour-ordered counter-based chain, directly: same accept decisions (a near-tie may flip one), states to FP32 rounding."""
    R, off, seed = 2, 11, 24680
    cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [random_symmetric_J(5)], 48, 1, n_replicas=R, precision=32, seed=seed, replica_offset=off, stream=stream)
    cfg.set_sweep_mode(mode)
    assert cfg.info()["tma"] and cfg.info()["tma_tc"] == (mode == 6)
    beta, sigma = np.array([2.0, 15.0]), np.array([0.5, 0.1])
    acc, att = cfg.sweep(1, beta, sigma, hits=2)
    for r in range(R):
        o = make_oracle(cfg, oracle_mod)
        o.randomize_v1(seed, off + r)
        a1 = o.sweeps_colour(stream, 1, 2, cfg.colour, beta[r], sigma[r], seed, off + r, 0)
        assert abs(int(acc[r]) - a1) <= 2, (acc[r], a1)
        d = np.abs(cfg.get_state(r, 1)[0] - o.get_state()).max(axis=1)          # per site
        assert (d > 1e-4).sum() <= 4, (d > 1e-4).sum()
        assert np.median(d) < 5e-6
    cfg.close()


@pytest.mark.parametrize("rc", [8, 64, 128])
def test_tma_replicas_per_cta_do_not_change_chains(scmc, monkeypatch, rc):
    """A CTA of the TMA-pipelined kernel walks over up to 128 replicas (alive replicas compacted 32 at a time, acceptance counts flushed every
    32 visits): whatever the chunking (ragged last chunk included), mode 5 realises the direct-gather chain bit for bit and mode 6 its own."""
    monkeypatch.setenv("SCMC_TMA_MINCTAS", "1")           # keep the requested chunk although the grid is small
    R = 150
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 48, 1, n_replicas=R, precision=32, seed=99, stream=2)
    beta = np.geomspace(0.5, 50.0, R); sigma = np.geomspace(0.8, 0.03, R)
    ref = mk(); ref.set_sweep_mode(3); ref.set_lanes(1)
    acc_r, _ = ref.sweep(2, beta, sigma, hits=2)
    monkeypatch.setenv("SCMC_TMA_RC", "16")
    tc_ref = mk(); tc_ref.set_sweep_mode(6); tc_ref.set_lanes(1)
    acc_t, _ = tc_ref.sweep(2, beta, sigma, hits=2)
    monkeypatch.setenv("SCMC_TMA_RC", str(rc))
    a = mk(); a.set_sweep_mode(5); a.set_lanes(1)
    acc_a, att_a = a.sweep(2, beta, sigma, hits=2)
    assert np.array_equal(acc_a, acc_r) and np.all(att_a == 2 * 2 * a.nSites) and np.array_equal(a.get_state(), ref.get_state())
    b = mk(); b.set_sweep_mode(6); b.set_lanes(1)
    acc_b, _ = b.sweep(2, beta, sigma, hits=2)
    assert np.array_equal(acc_b, acc_t) and np.array_equal(b.get_state(), tc_ref.get_state())
    for c in (ref, tc_ref, a, b):
        c.close()


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_more_than_eight_colours(scmc, oracle_mod, mode):
    """Up to 16 colour classes (the reference visits random sites, so any proper colouring is a legitimate order here): a 12-class refinement of
    the triangular 3-colouring; the FP64 chain of every sweep kernel follows the oracle's chain in the same class order."""
    seed = 2718
    base = scmc.initializeCfg("nearest-neighbor", "triangular", [random_symmetric_J(5)], 12, 1, n_replicas=1, precision=64, seed=seed)
    colour = base.colour + 3 * (np.arange(base.nSites) % 4)         # proper: neighbours already differ mod 3
    base.close()
    assert len(set(colour.tolist())) == 12
    cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [random_symmetric_J(5)], 12, 1, n_replicas=2, precision=64, seed=seed, replica_offset=3, colour=colour, stream=2)
    info = cfg.info()
    assert info["n_colours"] == 12
    if (mode == 2 and not info["resident"]) or (mode == 4 and not info["resident_psi"]):
        pytest.skip("resident kernel of this family not available for the model")
    cfg.set_sweep_mode(mode)
    o = make_oracle(cfg, oracle_mod)
    o.randomize_v1(seed, 4)
    acc, _ = cfg.sweep(3, np.array([1.0, 3.0]), np.array([0.5, 0.3]))
    a = o.sweeps_colour(2, 3, 2, cfg.colour, 3.0, 0.3, seed, 4, 0)
    assert acc[1] == a
    assert np.allclose(cfg.get_state(1, 1)[0], o.get_state(), atol=1e-10)
    assert abs(cfg.energies()[1] - o.energy()) < 1e-9 * cfg.nSites
    with pytest.raises(scmc._lib.ScmcError, match="colour"):
        scmc.initializeCfg("nearest-neighbor", "triangular", [random_symmetric_J(5)], 12, 1, n_replicas=1, precision=64, seed=seed,
                           colour=base.colour + 3 * (np.arange(base.nSites) % 6))      # 18 classes: over the limit


@pytest.mark.parametrize("n", [1, 3])
@pytest.mark.parametrize("precision", [32, 64])
def test_several_label_psi_gather_kernel(scmc, oracle_mod, n, precision):
    """TG/h-BN couplings at d = 4 (fillings 1, 3): the streaming psi-gather kernel for several coupling labels (k_sweep_colour_psi_nl: neighbours enter through
    their states, one density sum and one 16x16 mat-vec per label) is what auto runs on lattices too large for the resident T-gather kernel; mode 3 selects it
    anywhere.  Its chain equals the T-gather chain (mode 1) to rounding (FP64: 1e-9 after whole sweeps), energies / T vectors / measurement rows of the state it
    leaves agree with the oracle, and the lazily refreshed T cache is consistent."""
    mk = lambda R, L: scmc.initializeCfg("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.15, 0.8], L, n, n_replicas=R, precision=precision, seed=4711, stream=2)
    a = mk(3, 12); a.set_sweep_mode(1)
    b = mk(3, 12); b.set_sweep_mode(3)
    beta, sigma = np.array([1.0, 4.0, 20.0]), np.array([0.6, 0.3, 0.1])
    acc_a, _ = a.sweep(3, beta, sigma, hits=2)
    acc_b, _ = b.sweep(3, beta, sigma, hits=2)
    if precision == 64:
        assert np.array_equal(acc_a, acc_b)
        assert np.allclose(a.get_state(), b.get_state(), atol=1e-9)
    else:
        assert np.abs(acc_a.astype(np.int64) - acc_b.astype(np.int64)).max() <= 3
    tol = 1e-10 if precision == 64 else 2e-5
    for r in range(3):
        o = make_oracle(b, oracle_mod)
        o.set_state(b.get_state(r, 1)[0])
        assert abs(o.energy() - b.energies()[r]) < tol * b.nSites * 10
        assert np.allclose(b.get_T(r), o.get_T()[0], atol=tol * 10)          # T planes refreshed on demand after psi-gather sweeps
    rows = b.measure()
    assert np.allclose(rows[:, 0] * b.nSites, b.energies(), rtol=0, atol=tol * b.nSites * 10)
    # auto: large lattice / many replicas -> the several-label psi-gather kernel (launch count of a streaming call), same chain as mode 3
    c = mk(600, 48); d = mk(600, 48); d.set_sweep_mode(3)
    bt, sg = np.geomspace(0.5, 20.0, 600), np.geomspace(0.6, 0.05, 600)
    acc_c, _ = c.sweep(1, bt, sg, hits=2); acc_d, _ = d.sweep(1, bt, sg, hits=2)
    assert np.array_equal(acc_c, acc_d) and np.array_equal(c.get_state(), d.get_state())
    for x in (a, b, c, d):
        x.close()


def test_annealing_on_a_tiled_lattice_with_stopped_replicas(scmc):
    """runAnnealing!'s host-driven cooling loop on a lattice whose colour classes run through the TMA-pipelined kernel: replicas stop one by one (the `active`
    mask compacts the replica chunk of a CTA), so the kernels see ragged chunks of alive replicas.  Mode 5 = mode 3 bit for bit; the tensor-core variant
    (mode 6) reaches the same sweep counts within a few sweeps and the same energies within the spread of the restarts."""
    mk = lambda: scmc.initializeCfg("nearest-neighbor", "triangular", [NOTEBOOK_J], 48, 1, n_replicas=40, precision=32, seed=1357, stream=2)
    kw = dict(T_fac=0.8, N_max=260, N_o=0, N_per_T=8, min_acc_per_site=2, min_accrate=0.25, σ_min=0.05, log_rate=0)
    out = {}
    for mode in (3, 5, 6):
        c = mk(); c.set_sweep_mode(mode)
        assert c.info()["tma"]
        out[mode] = scmc.runAnnealing_(c, 2.0, **kw)
        out[mode]["state"] = c.get_state()
        c.close()
    a, b, t = out[3], out[5], out[6]
    assert len(set(a["sweeps"].tolist())) >= 2 and a["sweeps"].min() < 250         # replicas really stop at different times
    assert np.array_equal(a["sweeps"], b["sweeps"]) and np.array_equal(a["E"], b["E"]) and np.array_equal(a["state"], b["state"])
    assert np.abs(a["sweeps"] - t["sweeps"]).max() <= 2 * 8                          # at most a temperature step apart
    assert abs(np.mean(a["E"]) - np.mean(t["E"])) < 4 * np.std(a["E"]) / np.sqrt(40) + 1e-3 * abs(np.mean(a["E"]))


@pytest.mark.parametrize("case", [
    ("nearest-neighbor", "triangular", [NOTEBOOK_J], 128, 1),      # C5 lattice: seam classes + bulk-copy tile plan
    ("nearest-neighbor", "triangular", [NOTEBOOK_J], 48, 1),
    ("nearest-neighbor", "honeycomb", [NOTEBOOK_J], 24, 1),        # z = 3
    ("nearest-neighbor", "square", [NOTEBOOK_J], 32, 1),           # z = 4 padded to 6
    ("nearest-neighbor", "cubic", [NOTEBOOK_J], 5, 1),             # odd size, greedy colouring
    ("tg-hbn", "triangular", [1.0, 0.3, 0.2, 0.15, 0.8], 12, 2),   # 3 labels, 12 neighbours
])
def test_device_tables_self_check(scmc, monkeypatch, case):
    """scmc_verify_tables re-derives, from the DEVICE copies the kernels read, that no bond lies inside a colour class (the race-freedom condition of a
    colour pass), that bonds are symmetric, and that the class-position tables and the bulk-copy tile plan resolve to exactly the neighbour table.
    It finds nothing on the handles scmc_create returns and does find a planted inconsistency (self-test of the checker)."""
    model, lattice, J, L, n = case
    cfg = scmc.initializeCfg(model, lattice, J, L, n, n_replicas=2, precision=32, seed=1)
    assert cfg.verify_tables() == 0
    monkeypatch.setenv("SCMC_VERIFY_PLANT", "1")
    assert cfg.verify_tables() >= 2
    monkeypatch.delenv("SCMC_VERIFY_PLANT")
    assert cfg.verify_tables() == 0                                  # the planted entry was restored
    acc, att = cfg.sweep(1, np.array([1.0, 2.0]), np.array([0.3, 0.3]))
    assert np.all(acc > 0)
    cfg.close()
