"""Dev tool: the factored-model probs kernel on the bench's C5 circuits (for ncu)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygsti_b200 import engine, fixtures as fx
from pygsti_b200.packing import FactoredModel
import torch
ctx = engine.Context(0)
G, rho, E = fx.random_dense_model(256, 14, 1, 16, seed=1)
t, _ = fx.random_layout(256, 14, 16, 5000, 128, seed=0)
rng = np.random.default_rng(5)
fptr, f_nq, f_t, f_off, mats, off = [0], [], [], [], [], 0
targets = [(q,) for q in range(4)] * 2 + [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2)]
for tg in targets:
    for _ in range(2):
        k = len(tg); small = np.eye(4 ** k) * 0.9 + 0.2 * rng.standard_normal((4 ** k, 4 ** k)) / 2 ** k
        f_nq.append(k); f_t.append(list(tg) + [-1] * (4 - k)); f_off.append(off); mats.append(small.ravel()); off += small.size
    fptr.append(len(f_nq))
fm = FactoredModel(n_qubits=4, op_fptr=np.asarray(fptr, np.int32), f_nq=np.asarray(f_nq, np.int32), f_targets=np.asarray(f_t, np.int32).reshape(-1, 4),
                   f_moff=np.asarray(f_off, np.int64), mats=np.concatenate(mats), rho=rho, E=E)
at = ctx.upload_atom(t); at.set_model_factored(fm)
P = torch.empty(t.n_elements, dtype=torch.float64, device="cuda")
for _ in range(3):
    t0 = time.time(); at.fill_probs_dev(P.data_ptr()); ctx.sync(); print("factored probs ms", (time.time() - t0) * 1e3)
