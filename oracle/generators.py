"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the reference's su(4) generator construction,
`/root/reference/src/Generators.jl:2-65`.  It follows the reference step by
step on purpose (occupation-number vectors, parton matrices `f`, full-Fock-space
operator, projection onto filling n), including the reference's quirk that the
parton operator carries NO Jordan-Wigner sign (`Generators.jl:47`).

Parity status: the reference has no golden vectors for this (empty test suite,
`test/runtests.jl:4-6`); pinned by the analytic identities in
`tests/test_oracle_generators.py` (kron structure for n=1, spectra / nnz for n=2).
"""
import numpy as np


def pauli(mu, normalization=1.0):
    """sigma^mu, mu=0 identity (Generators.jl:68-80)."""
    mats = np.zeros((4, 2, 2), dtype=np.complex128)
    mats[0] = [[1, 0], [0, 1]]
    mats[1] = [[0, 1], [1, 0]]
    mats[2] = [[0, -1j], [1j, 0]]
    mats[3] = [[1, 0], [0, -1]]
    mats[1:] /= normalization
    return mats[mu]


def sl_to_j(s, l):
    """(spin s in 1..2, valley l in 1..2) -> mode 1..4 (Generators.jl:34-36)."""
    return 2 * s + l - 2


def fock_basis():
    """Occupation vectors of integers 0..15, little-endian (Generators.jl:4)."""
    return [[(m >> b) & 1 for b in range(4)] for m in range(16)]


def parton(s, l, basis):
    """<a_i| f_{sl} |a_j> without fermionic sign (Generators.jl:39-51)."""
    nb = len(basis)
    mat = np.zeros((nb, nb))
    mode = sl_to_j(s, l) - 1
    for j in range(nb):
        lowered = list(basis[j])
        lowered[mode] -= 1
        for i in range(nb):
            if basis[i] == lowered:
                mat[i, j] = 1.0
                break
    return mat


def spin_full(mu, nu, basis):
    """sum_{s s' l l'} f_{sl}^dagger f_{s'l'} sigma^mu_{ss'} sigma^nu_{ll'} (Generators.jl:54-56)."""
    nb = len(basis)
    out = np.zeros((nb, nb), dtype=np.complex128)
    smu, snu = pauli(mu), pauli(nu)
    for s in (1, 2):
        for sp in (1, 2):
            for l in (1, 2):
                for lp in (1, 2):
                    out += (parton(s, l, basis).T @ parton(sp, lp, basis)) * smu[s - 1, sp - 1] * snu[l - 1, lp - 1]
    return out


def project(op, n, basis):
    """Restrict to the n-parton sector (Generators.jl:59-62)."""
    keep = [i for i, b in enumerate(basis) if sum(b) == n]
    return op[np.ix_(keep, keep)]


def get_generators(n):
    """16 d x d matrices, a = 4 mu + nu (0-based) (Generators.jl:2-28)."""
    basis = fock_basis()
    return np.stack([project(spin_full(mu, nu, basis), n, basis) for mu in range(4) for nu in range(4)])


def get_generators_sq(n):
    """Matrix squares (`gen .^ 2` on a Vector{Matrix}, Configuration.jl:52)."""
    g = get_generators(n)
    return np.stack([m @ m for m in g])
