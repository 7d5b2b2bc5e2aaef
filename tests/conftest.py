import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


class Case:
    """A golden fixture: layout tables + model tensors + reference outputs (tests/golden/make_golden.py)."""

    def __init__(self, name):
        from pygsti_b200.packing import AtomTables, DerivMap
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        z = self.z
        self.n_atoms = int(z["n_atoms"])
        self.n_elements = int(z["n_elements"])
        self.num_params = int(z["num_params"])
        self.dim = int(z["dim"])
        self.atoms = []
        for i in range(self.n_atoms):
            pre = "a%d_" % i
            t = AtomTables.from_dict(z, pre)
            D = DerivMap(int(z[pre + "D_shape"][0]), int(z[pre + "D_shape"][1]),
                         z[pre + "D_rows"], z[pre + "D_cols"], z[pre + "D_vals"])
            es = z[pre + "element_slice"]
            self.atoms.append(dict(tables=t, G=z[pre + "G"], rho=z[pre + "rho"], E=z[pre + "E"], D=D,
                                   element_slice=slice(int(es[0]), int(es[1]))))

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z


@pytest.fixture(scope="session")
def load_case():
    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = Case(name)
        return cache[name]
    return _load


@pytest.fixture(scope="session")
def gpu_ctx():
    from pygsti_b200 import engine
    ctx = engine.Context(0)
    yield ctx
    ctx.close()
