import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


NOTEBOOK_J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)  # Is=-1, Iv=Isv=0.2 (examples.ipynb cells 4/7)


def random_symmetric_J(seed, zero11=True):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (16, 16))
    J = 0.5 * (A + A.T)
    if zero11:
        J[0, 0] = 0.0
    return J


def random_states(rng, *shape_d):
    v = rng.standard_normal(shape_d) + 1j * rng.standard_normal(shape_d)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


@pytest.fixture(scope="session")
def scmc():
    import scmc_b200
    return scmc_b200


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


def make_oracle(cfg, oracle_mod, replica=None):
    """An oracle chain with the same tables as a product Configuration (and optionally the state of one replica)."""
    o = oracle_mod.Oracle(cfg.interactionSites, cfg.interactionLabels, np.stack(cfg.interactions), cfg.onsiteInteraction, cfg.n,
                          generators=cfg.generators)
    if replica is not None:
        o.set_state(cfg.get_state(replica, 1)[0])
    return o


@pytest.fixture(scope="session")
def golden():
    """Recorded outputs of the reference's notebook (tests/golden/make_golden.py)."""
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "notebook_outputs.json")))
