"""End-to-end (host buffers, C ABI) timing of one C2 Jacobian for several B200_D2H_SPLIT settings (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case
c = Case("c2_full_layout"); a = c.atoms[0]
ctx = engine.Context(0)
at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
nE, Np = c.n_elements, c.num_params
Jh = engine.pinned_empty((nE, Np)); Ph = engine.pinned_empty((nE,))
for split in (sys.argv[1:] or ["1", "2", "3", "4", "1", "2"]):
    os.environ["B200_D2H_SPLIT"] = split
    ts = []
    for r in range(6):
        t0 = time.time(); at.set_model(a["G"], a["rho"], a["E"]); at.fill_dprobs(Jh, Ph); ctx.sync(); ts.append(time.time() - t0)
    ms = 1e3 * min(ts[1:])
    print("B200_D2H_SPLIT=%s: e2e %.2f ms (median %.2f) -> %.1f GB/s, %.3e outcomes/s" % (split, ms, 1e3 * float(np.median(ts[1:])), nE * (Np + 1) * 8 / ms / 1e6, nE / ms * 1e3), flush=True)
rows = c["dprobs_matrix_sample_elements"]
print("err", float(np.max(np.abs(Jh[rows] - c["dprobs_matrix_sample_rows"]))))
