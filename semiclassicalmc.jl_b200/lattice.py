"""Minimal replacement for the LatticePhysics.jl calls the hot path depends on.

The reference only uses LatticePhysics at construction time
(/root/reference/src/Configuration.jl:37-48, src/Models.jl:42,93-110,
src/Observables/Observables.jl:91-98): `getUnitcell`, `addBond!`,
`getLatticePeriodic`, `organizedBondsFrom`, `getReciprocalUnitcell`,
`getLatticeInBox`.  LatticePhysics is an unvendored, unpinned dependency
(Project.toml:9-16), so site ordering and bond order here are this repo's own
convention ("parity unpinned"): energies depend only on topology and
bond labels, which cross the C ABI as explicit tables.

Conventions: lattice vectors a1 = (sqrt3/2, -1/2), a2 = (sqrt3/2, +1/2); site
index (0-based) = b + n_basis * (j + L * i) for cell (i, j) and basis site b;
position = i a1 + j a2 + basis[b].
"""
from dataclasses import dataclass, field
import numpy as np

_S3 = np.sqrt(3.0)


@dataclass
class Bond:
    src: int          # basis site the bond starts from (0-based)
    dst: int          # basis site it ends on (0-based)
    label: int        # 1-based coupling label, as in the reference
    offset: tuple     # cell offset (o1, o2, ...)


@dataclass
class Unitcell:
    lattice_vectors: list
    sites: list                      # basis positions
    bonds: list = field(default_factory=list)

    @property
    def dim(self):
        return len(self.lattice_vectors)


def addBond(uc, src, dst, label, offset, both_directions=False):
    """`addBond!(uc, from, to, label, offset, both)`; src/dst are 1-based like the reference (Models.jl:98-110)."""
    offset = tuple(int(o) for o in offset)
    uc.bonds.append(Bond(src - 1, dst - 1, label, offset))
    if both_directions:
        uc.bonds.append(Bond(dst - 1, src - 1, label, tuple(-o for o in offset)))


def getUnitcell(name):
    """Unit cells with their default nearest-neighbour bonds (all label 1)."""
    name = str(name)
    a1 = np.array([_S3 / 2, -0.5])
    a2 = np.array([_S3 / 2, 0.5])
    if name == "triangular":
        uc = Unitcell([a1, a2], [np.zeros(2)])
        for off in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, -1), (-1, 1)):
            addBond(uc, 1, 1, 1, off)
    elif name == "honeycomb":
        uc = Unitcell([a1, a2], [np.zeros(2), np.array([1 / _S3, 0.0])])
        for off in ((0, 0), (-1, 0), (0, -1)):
            addBond(uc, 1, 2, 1, off)
        for off in ((0, 0), (1, 0), (0, 1)):
            addBond(uc, 2, 1, 1, off)
    elif name == "square":
        uc = Unitcell([np.array([1.0, 0.0]), np.array([0.0, 1.0])], [np.zeros(2)])
        for off in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            addBond(uc, 1, 1, 1, off)
    elif name == "cubic":
        uc = Unitcell([np.eye(3)[0], np.eye(3)[1], np.eye(3)[2]], [np.zeros(3)])
        for ax in range(3):
            for sgn in (1, -1):
                off = [0, 0, 0]
                off[ax] = sgn
                addBond(uc, 1, 1, 1, off)
    elif name == "chain":
        uc = Unitcell([np.array([1.0])], [np.zeros(1)])
        addBond(uc, 1, 1, 1, (1,))
        addBond(uc, 1, 1, 1, (-1,))
    else:
        raise ValueError(f"unit cell '{name}' not available")
    return uc


@dataclass
class Lattice:
    unitcell: Unitcell
    L: int
    positions: np.ndarray      # [N, dim]
    basis_labels: np.ndarray   # [N] 1-based like the reference's `label.(l.sites)`
    cells: np.ndarray          # [N, dim] integer cell coordinates
    nbr_site: np.ndarray       # [N, z] 0-based neighbour site, -1 = padding
    nbr_label: np.ndarray      # [N, z] 1-based coupling label, 0 = padding

    @property
    def n_sites(self):
        return self.positions.shape[0]


def getLatticePeriodic(uc, L):
    """L^dim unit cells with periodic boundaries; bonds organised per source site (`organizedBondsFrom`)."""
    dim, nb = uc.dim, len(uc.sites)
    N = nb * L ** dim
    grids = np.stack(np.meshgrid(*[np.arange(L)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)  # i slowest
    cells = np.repeat(grids, nb, axis=0)
    bidx = np.tile(np.arange(nb), L ** dim)
    vecs = np.stack(uc.lattice_vectors)
    positions = cells @ vecs + np.stack(uc.sites)[bidx]
    strides = np.array([L ** (dim - 1 - k) for k in range(dim)])
    per_basis = [[b for b in uc.bonds if b.src == s] for s in range(nb)]
    z = max(len(p) for p in per_basis)
    nbr_site = -np.ones((N, z), dtype=np.int32)
    nbr_label = np.zeros((N, z), dtype=np.int32)
    for s in range(nb):
        mine = np.nonzero(bidx == s)[0]
        for k, b in enumerate(per_basis[s]):
            tgt = (cells[mine] + np.array(b.offset)) % L
            nbr_site[mine, k] = b.dst + nb * (tgt @ strides)
            nbr_label[mine, k] = b.label
    return Lattice(uc, L, positions, bidx + 1, cells, nbr_site, nbr_label)


def getReciprocalVectors(lattice_vectors):
    """b_i with a_i . b_j = 2 pi delta_ij."""
    A = np.stack(lattice_vectors)
    return 2 * np.pi * np.linalg.inv(A).T


def getKsInBox(lattice, L, dimensions, center):
    """Allowed momenta of the L-supercell inside a box (Observables.jl:91-98)."""
    uc = getUnitcell(lattice)
    B = getReciprocalVectors([L * v for v in uc.lattice_vectors])
    dimensions = np.asarray(dimensions, dtype=float)
    center = np.asarray(center, dtype=float)
    reach = int(np.ceil(np.linalg.norm(dimensions / 2 + np.abs(center)) / np.min(np.linalg.norm(B, axis=1)))) * 2 + 2
    rng = np.arange(-reach, reach + 1)
    mesh = np.stack(np.meshgrid(*[rng] * len(B), indexing="ij"), axis=-1).reshape(-1, len(B))
    ks = mesh @ B
    keep = np.all(np.abs(ks - center) <= dimensions / 2 + 1e-9, axis=1)
    return ks[keep]


def colourSites(lat, prefer_few_passes=True):
    """Graph colouring of the interaction graph (no two bonded sites share a colour).

    Closed forms where they exist (fewest colour passes), else greedy:
    bipartite honeycomb/square/cubic/chain nn -> 2; triangular nn with L % 3 == 0 -> 3;
    triangular (nn or nn+nnn) with even L -> 4 (2x2 cell parity).
    """
    N = lat.n_sites
    cands, fallback = [], []
    c = lat.cells
    nb = len(lat.unitcell.sites)
    if nb == 2:
        cands.append(lat.basis_labels - 1)
    if lat.L % 2 == 0:
        cands.append(c.sum(axis=1) % 2)
    if c.shape[1] == 2:
        if lat.L % 3 == 0:
            cands.append((c[:, 0] - c[:, 1]) % 3)
        if lat.L % 2 == 0:
            fallback.append(((c[:, 0] % 2) + 2 * (c[:, 1] % 2)).astype(np.int32))
    for col in cands:
        col = np.asarray(col, dtype=np.int32)
        if _valid_colouring(lat.nbr_site, col):
            return col
    if c.shape[1] == 2 and nb == 1 and lat.L >= 6 and prefer_few_passes:
        # triangular-type lattice with L % 3 != 0: the 3-colouring (x - y) % 3 only fails on bonds that cross the periodic
        # seam; recolouring the O(L) seam sites gives 3 large classes + a few tiny ones.  Every T vector is then read by
        # 2 instead of 3 other colour passes per sweep (less HBM traffic than the 4-colour 2x2 scheme).
        col = repairColouring(lat.nbr_site, ((c[:, 0] - c[:, 1]) % 3).astype(np.int32))
        if col is not None:
            return col
    for col in fallback:
        if _valid_colouring(lat.nbr_site, col):
            return col
    return greedyColouring(lat.nbr_site)


def repairColouring(nbr_site, colour, max_colours=8):
    """Make `colour` a valid colouring by moving one endpoint of every conflicting bond to an extra colour class."""
    colour = np.array(colour, dtype=np.int32)
    N, z = nbr_site.shape
    adj = [set() for _ in range(N)]
    for k in range(z):
        for i in np.nonzero(nbr_site[:, k] >= 0)[0]:
            j = int(nbr_site[i, k])
            adj[i].add(j); adj[j].add(int(i))
    base = int(colour.max()) + 1
    for i in range(N):
        if any(colour[j] == colour[i] for j in adj[i] if j > i):
            used = {int(colour[j]) for j in adj[i]}
            cnew = base
            while cnew in used:
                cnew += 1
            if cnew >= max_colours:
                return None
            colour[i] = cnew
    if np.count_nonzero(colour >= base) > 0.1 * N:      # not a seam repair any more: let the caller use another scheme
        return None
    return colour if _valid_colouring(nbr_site, colour) else None


def _valid_colouring(nbr_site, colour):
    for k in range(nbr_site.shape[1]):
        j = nbr_site[:, k]
        ok = j >= 0
        if np.any(colour[ok] == colour[j[ok]]):
            return False
    return True


def greedyColouring(nbr_site):
    N, z = nbr_site.shape
    adj = [set() for _ in range(N)]
    for i in range(N):
        for j in nbr_site[i]:
            if j >= 0:
                adj[i].add(int(j))
                adj[int(j)].add(i)
    colour = -np.ones(N, dtype=np.int32)
    for i in range(N):
        used = {colour[j] for j in adj[i] if colour[j] >= 0}
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    return colour
