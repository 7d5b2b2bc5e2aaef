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
            if pre + "F_op_fptr" in z:           # gates also as factor programs + the factor-space derivative map
                from .packing import FactoredModel
                self.atoms[-1]["fm"] = FactoredModel(n_qubits=int(z[pre + "F_n_qubits"]), op_fptr=z[pre + "F_op_fptr"], f_nq=z[pre + "F_nq"],
                                                     f_targets=z[pre + "F_targets"], f_moff=z[pre + "F_moff"], mats=z[pre + "F_mats"],
                                                     rho=z[pre + "rho"], E=z[pre + "E"])
                self.atoms[-1]["Df"] = DerivMap(int(z[pre + "FD_shape"][0]), int(z[pre + "FD_shape"][1]),
                                                z[pre + "FD_rows"], z[pre + "FD_cols"], z[pre + "FD_vals"])

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


# ---- synthetic layouts of BASELINE configs 3 and 5 (rng seed 0, depth ~ U{1..max_depth}, every layer drawn
# uniformly from the layer labels, every circuit measured in all effects) ------------------------------------------------
def random_layout(dim, n_ops, n_eff, n_circuits, max_depth, seed=0, rows=None):
    """(AtomTables, ops list) of `n_circuits` independent random circuits (no prefix sharing: cache_size 0); element
    index = circuit * n_eff + effect.  `rows` = (lo, hi) keeps only that block of circuits (a shard), with local elements
    renumbered from 0 -- the random stream is the same whatever the block, so shards of different ranks fit together."""
    rng = np.random.default_rng(seed)
    circs = [rng.integers(0, n_ops, size=int(rng.integers(1, max_depth + 1))).astype(np.int32) for _ in range(n_circuits)]
    lo, hi = (0, n_circuits) if rows is None else rows
    sel = circs[lo:hi]
    n = len(sel)
    ptr = np.zeros(n + 1, np.int64)
    if n:
        ptr[1:] = np.cumsum([len(c) for c in sel])
    t = AtomTables(dim=dim, n_ops=n_ops, n_rho=1, n_eff=n_eff, n_elements=n * n_eff, cache_size=0,
                   row_dest=np.arange(n, dtype=np.int32), row_istart=np.full(n, -1, np.int32),
                   row_icache=np.full(n, -1, np.int32), row_prep=np.zeros(n, np.int32),
                   row_ptr=ptr.astype(np.int32),
                   row_ops=(np.concatenate(sel).astype(np.int32) if n else np.zeros(0, np.int32)),
                   out_ptr=(np.arange(n + 1, dtype=np.int64) * n_eff).astype(np.int32),
                   out_eff=np.tile(np.arange(n_eff, dtype=np.int32), n),
                   out_el=np.arange(n * n_eff, dtype=np.int32))
    return t, circs


def balanced_blocks(circs, world, n_eff, store_weight=8.0):
    """Cut a list of circuits into `world` contiguous blocks of equal estimated work (depth x (1 + outcomes) sweeps
    + the Jacobian rows written); returns [(lo, hi)] per rank."""
    w = np.array([(len(c) + 1) * (1.0 + n_eff) + store_weight * n_eff for c in circs], dtype=np.float64)
    cw = np.cumsum(w)
    cuts = [0] + [int(np.searchsorted(cw, cw[-1] * r / world)) for r in range(1, world)] + [len(circs)]
    cuts = np.maximum.accumulate(np.asarray(cuts))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def random_dense_model(dim, n_ops, n_rho, n_eff, seed=0):
    """Contractive random dense superoperators / superkets / effects (BASELINE config 5 stand-in: d = 256, 14 layer labels)."""
    rng = np.random.default_rng(seed)
    G = np.eye(dim)[None] * 0.9 + 0.3 / np.sqrt(dim) * rng.standard_normal((n_ops, dim, dim))
    rho = rng.standard_normal((n_rho, dim)) / np.sqrt(dim)
    E = rng.standard_normal((n_eff, dim)) / np.sqrt(dim)
    return G, rho, E
