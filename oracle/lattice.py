"""ORACLE (test infrastructure only -- never imported by the product path).

Independent construction of the periodic triangular / honeycomb neighbour tables
the reference obtains from LatticePhysics.jl (`getLatticePeriodic`,
`organizedBondsFrom`; call sites /root/reference/src/Configuration.jl:37-48,
src/Models.jl:42,93-110).  LatticePhysics is not vendored and no version is
pinned (Project.toml:9-16): **parity unpinned** for site ordering; the physics
only depends on topology + labels.  Plain loops, no shared code with the product.
"""
import numpy as np

S3 = 3.0 ** 0.5
A1 = (S3 / 2, -0.5)
A2 = (S3 / 2, 0.5)

TRI_NN = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, -1), (-1, 1)]
# tg-hbn bond set (Models.jl:98-110): label 1, label 2 (reversed directions), label 3 nnn both ways
TGHBN = [((1, 0), 1), ((0, -1), 1), ((-1, 1), 1),
         ((-1, 0), 2), ((0, 1), 2), ((1, -1), 2),
         ((1, 1), 3), ((-1, -1), 3), ((-1, 2), 3), ((1, -2), 3), ((-2, 1), 3), ((2, -1), 3)]


def triangular(L, bonds=None):
    """Returns (positions [N,2], nbr_site [N,z], nbr_label [N,z] (1-based), basis_label [N])."""
    bonds = [(o, 1) for o in TRI_NN] if bonds is None else bonds
    N = L * L
    pos = np.zeros((N, 2))
    nbr = np.zeros((N, len(bonds)), dtype=np.int32)
    lab = np.zeros((N, len(bonds)), dtype=np.int32)
    for i in range(L):
        for j in range(L):
            s = j + L * i
            pos[s] = (i * A1[0] + j * A2[0], i * A1[1] + j * A2[1])
            for k, ((o1, o2), label) in enumerate(bonds):
                nbr[s, k] = ((j + o2) % L) + L * ((i + o1) % L)
                lab[s, k] = label
    return pos, nbr, lab, np.ones(N, dtype=np.int64)


def honeycomb(L):
    N = 2 * L * L
    pos = np.zeros((N, 2))
    nbr = np.zeros((N, 3), dtype=np.int32)
    lab = np.ones((N, 3), dtype=np.int32)
    bl = np.zeros(N, dtype=np.int64)
    from_a = [(0, 0), (-1, 0), (0, -1)]
    from_b = [(0, 0), (1, 0), (0, 1)]
    for i in range(L):
        for j in range(L):
            a = 2 * (j + L * i)
            b = a + 1
            x, y = i * A1[0] + j * A2[0], i * A1[1] + j * A2[1]
            pos[a] = (x, y)
            pos[b] = (x + 1 / S3, y)
            bl[a], bl[b] = 1, 2
            for k, (o1, o2) in enumerate(from_a):
                nbr[a, k] = 2 * (((j + o2) % L) + L * ((i + o1) % L)) + 1
            for k, (o1, o2) in enumerate(from_b):
                nbr[b, k] = 2 * (((j + o2) % L) + L * ((i + o1) % L))
    return pos, nbr, lab, bl


def basis_vectors(name):
    return [np.zeros(2)] if name == "triangular" else [np.zeros(2), np.array([1 / S3, 0.0])]


def unit_vectors():
    return [np.array(A1), np.array(A2)]
