"""Pins the oracle's generator restatement (oracle/generators.py <- Generators.jl:2-65) with the analytic identities of SURVEY section 4,
and the product's independent construction + generated CUDA tables against it."""
import numpy as np

from oracle import generators as og

PAULI = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])]


def test_n1_is_kron_of_paulis():
    g = og.get_generators(1)
    assert g.shape == (16, 4, 4)
    for mu in range(4):
        for nu in range(4):
            assert np.allclose(g[4 * mu + nu], np.kron(PAULI[mu], PAULI[nu]))
    assert np.allclose(og.get_generators_sq(1), np.eye(4)[None])


def test_n2_reference_conventions():
    g = og.get_generators(2)
    assert g.shape == (16, 6, 6)
    assert np.allclose(g[0], 2 * np.eye(6))
    assert [int((np.abs(m) > 0).sum()) for m in g] == [6, 8, 8, 2, 8, 8, 8, 8, 8, 8, 8, 8, 2, 8, 8, 2]
    for a in range(1, 16):
        assert np.allclose(g[a], g[a].conj().T)
        assert np.allclose(np.sort(np.linalg.eigvalsh(g[a])), [-2, 0, 0, 0, 0, 2])
    assert np.allclose(sum(m @ m for m in g[1:]), 20 * np.eye(6))
    vals = np.unique(np.round(np.concatenate([g.real.ravel(), g.imag.ravel()]), 12))
    assert set(vals.tolist()) <= {-2.0, -1.0, 0.0, 1.0, 2.0}


def test_n2_signless_quirk_breaks_algebra():
    """without Jordan-Wigner signs the n=2 matrices do not close under commutation."""
    g = og.get_generators(2)
    basis = g.reshape(16, -1)
    outside = 0
    for a in range(16):
        for b in range(16):
            c = (g[a] @ g[b] - g[b] @ g[a]).reshape(-1)
            coef, *_ = np.linalg.lstsq(basis.T, c, rcond=None)
            outside += np.linalg.norm(basis.T @ coef - c) > 1e-9
    assert outside > 0


def test_product_generators_equal_oracle(scmc):
    for n in (1, 2, 3):
        assert np.array_equal(scmc.generators.getGenerators(n), og.get_generators(n))
        assert np.allclose(scmc.generators.getGeneratorsSq(n), og.get_generators_sq(n))
    assert scmc.get_su4index(2, 3) == 12


def test_generated_cuda_tables_match(tmp_path, scmc):
    """csrc/gen_code.h (committed) is what the generator emits today, and its integer tables equal the oracle's matrices."""
    import os, re
    out = tmp_path / "gen_code.h"
    scmc.generators.emit_cuda_tables(str(out))
    committed = os.path.join(os.path.dirname(scmc.__file__), "csrc", "gen_code.h")
    text = open(committed).read()
    assert out.read_text() == text
    for n in (1, 2, 3):
        g = og.get_generators(n)
        for part, arr in (("re", g.real), ("im", g.imag)):
            m = re.search(rf"G{n}_{part}\[[^\]]*\] = \{{([^}}]*)\}}", text)
            tab = np.array([int(x) for x in m.group(1).split(",")]).reshape(g.shape)
            assert np.array_equal(tab, arr)


def _run_generated(text, name, **arrays):
    """Execute the body of the generated (scalar, template) CUDA function `name` in Python: the bodies are straight-line arithmetic on
    re[], im[], h[], Hq[] with fma(a, b, c) and R(0)."""
    import re as _re
    m = _re.search(r"template <class R>[^\n]*\b" + name + r"\(([^)]*)\) \{\n(.*?)\n\}\n", text, _re.S)
    assert m, name
    env = dict(arrays, fma=lambda a, b, c: a * b + c, R=float)
    ret = None
    for line in m.group(2).splitlines():
        for stmt in line.split(";"):
            stmt = stmt.split("//")[0].strip()
            if not stmt:
                continue
            if stmt.startswith("return "):
                ret = eval(stmt[7:], env)
                continue
            stmt = _re.sub(r"^(const )?R ", "", stmt)
            if "," in stmt and "fma(" not in stmt and "=" in stmt and stmt.count("=") > 1:      # "ea = R(0), eb = R(0)"
                for part in stmt.split(","):
                    exec(part.strip(), env)
            else:
                exec(stmt, env)
    return ret


def test_generated_energy_row_form_equals_the_field_contraction(scmc):
    """The generated local energy of the psi-gather family, e(psi) = psi^dagger (H psi) evaluated row by row over the upper triangle
    (energy_from_state_g*, csrc/gen_code.h), against its definition e = sum_a h^a <psi|G^a|psi> (exchangeEnergy, src/Configuration.jl:292-296, with the
    neighbour mat-vecs hoisted into h) for n = 1, 2, 3 -- including n = 2, where only 30 of the 36 density entries occur and rows have gaps."""
    import os
    text = open(os.path.join(os.path.dirname(scmc.__file__), "csrc", "gen_code.h")).read()
    rng = np.random.default_rng(12)
    for n in (1, 2, 3):
        g = og.get_generators(n)
        d = g.shape[1]
        for _ in range(5):
            psi = rng.standard_normal(d) + 1j * rng.standard_normal(d)
            psi /= np.linalg.norm(psi)
            h = rng.standard_normal(16)
            T = np.einsum("j,ajk,k->a", psi.conj(), g, psi).real
            nd = int(re_search_int(text, rf"NDENS_g{n} = (\d+)"))
            Hq = [0.0] * nd
            _run_generated(text, f"hq_from_field_g{n}", h=list(h), Hq=Hq)
            e = _run_generated(text, f"energy_from_state_g{n}", re=list(psi.real), im=list(psi.imag), Hq=Hq)
            assert abs(e - float(h @ T)) < 1e-12 * max(1.0, abs(h @ T)), (n, e, h @ T)
            # the density basis itself: P = density_acc(psi), T = expect_from_density(P), e = Hq . P
            P = [0.0] * nd
            _run_generated(text, f"density_acc_g{n}", re=list(psi.real), im=list(psi.imag), P=P)
            T2 = [0.0] * 16
            _run_generated(text, f"expect_from_density_g{n}", P=P, T=T2)
            assert np.allclose(T2, T, atol=1e-13) and abs(float(np.dot(Hq, P)) - e) < 1e-12


def re_search_int(text, pattern):
    import re as _re
    return _re.search(pattern, text).group(1)
