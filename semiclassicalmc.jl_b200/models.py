"""Model construction -- host mirror of /root/reference/src/Models.jl (bond / coupling tables only; geometry from lattice.py)."""
import logging

import numpy as np

from .configuration import Configuration
from .lattice import getUnitcell, addBond


def initializeCfg(model, latticename, J, L, n, **kw):
    """initializeCfg (Models.jl:2-17).  Unknown model: logs an error and returns None like the reference (no throw)."""
    if model == "nearest-neighbor":
        return initializeCfgNN(J, latticename, L, n, **kw)
    if model == "tg-hbn":
        return initializeCfgTgHbn(J, L, n=n, **kw)
    logging.error("model %s not implemented", model)
    return None


def initializeCfgNN(J, latticename, L, n, **kw):
    """One coupling matrix on the default nn bonds, optional diagonal on-site term (Models.jl:23-45)."""
    J = [np.asarray(m, dtype=float) for m in J]
    if len(J) == 1:
        onsite = np.zeros(16)
        interactions = [J[0]]
    elif len(J) == 2:
        onsite = np.diag(J[0]).copy()
        interactions = [J[1]]
    else:
        raise ValueError("Need length(J) = 1 (nn interactions) or length(J) = 2 (on-site and nearest-neighbor interactions).")
    uc = getUnitcell(latticename)
    return Configuration(latticename, L, uc, interactions, onsite, n, **kw)


def initializeCfgTgHbn(J, L, n=2, **kw):
    """TG/h-BN model of PRB 99, 205150 (Models.jl:48-113): labels 1/2 = nn with +/- DM valley coupling, 3 = nnn."""
    assert len(J) == 5, "Need J = [J1, J2, Jp1, Jp2, JH]"
    J1, J2, Jp1, Jp2, JH = (np.asarray(J, dtype=float) / 8)
    onsite = np.array([0.0, 0.0, 0.0, JH / 2] + [-JH / 2, 0.0, 0.0, JH / 2] * 3)
    Js = np.eye(4)
    Jv1 = np.array([[J1, 0, 0, 0], [0, J1 + Jp1, Jp2, 0], [0, -Jp2, J1 + Jp1, 0], [0, 0, 0, J1]], dtype=float)
    Jv2 = Jv1.T.copy()
    J_nn1 = np.kron(Js, Jv1); J_nn1[0, 0] = 0.0
    J_nn2 = np.kron(Js, Jv2); J_nn2[0, 0] = 0.0
    J_nnn = np.diag([0.0] + [J2] * 15)
    uc = getUnitcell("triangular")
    uc.bonds = []
    for off in ((1, 0), (0, -1), (-1, 1)):
        addBond(uc, 1, 1, 1, off, False)
    for off in ((-1, 0), (0, 1), (1, -1)):
        addBond(uc, 1, 1, 2, off, False)
    for off in ((1, 1), (-1, 2), (-2, 1)):
        addBond(uc, 1, 1, 3, off, True)
    return Configuration("triangular", L, uc, [J_nn1, J_nn2, J_nnn], onsite, n, **kw)
