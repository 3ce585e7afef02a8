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
