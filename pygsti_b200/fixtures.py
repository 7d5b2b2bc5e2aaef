"""Loader for packed layout files (.npz): integer atom tables + dense model tensors + derivative map
(+ optional reference outputs).  The format is what ``tests/golden/make_golden.py`` writes with
``pygsti_b200.packing``; ``bench.py`` and the tests read their workloads through this module so that
nothing needs pyGSTi at run time."""
import os

import numpy as np

from .packing import AtomTables, DerivMap

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


class Case:
    def __init__(self, name_or_path):
        path = name_or_path if os.path.isabs(name_or_path) or name_or_path.endswith(".npz") \
            else os.path.join(GOLDEN_DIR, name_or_path + ".npz")
        self.name = os.path.splitext(os.path.basename(path))[0]
        self.z = np.load(path)
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

    def hess_map(self, tag="H2", atom=0):
        """Second-derivative map stored by make_golden.py (`a<atom>_<tag>_*`): full Hessian ("H2") or rectangle i
        ("H2r<i>")."""
        from .packing import HessMap
        z, pre = self.z, "a%d_%s_" % (atom, tag)
        sh = z[pre + "shape"]
        return HessMap(int(sh[0]), int(sh[1]), int(sh[2]), z[pre + "rows"], z[pre + "a"], z[pre + "b"], z[pre + "vals"])

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z
