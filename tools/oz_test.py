"""Ozaki (tcgen05 int8) J^T J vs the FP64 DMMA SYRK on a golden layout: parity + timing (dev tool, not the bench)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case

name = sys.argv[1] if len(sys.argv) > 1 else "c2_full_layout"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = Case(name)
a = c.atoms[0]
ctx = engine.Context(0)
at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
nE, Np = c.n_elements, c.num_params
rng = np.random.default_rng(0)
w = rng.uniform(0.5, 1.5, nE); f = rng.standard_normal(nE)
res = {}
for mode in ("dmma", "ozaki", "ozaki7"):
    ctx.set_jtj_mode({"dmma": 0, "ozaki": 8, "ozaki7": 7}[mode])
    ts = []
    try:
        for r in range(reps):
            t0 = time.time(); JTJ, JTf = at.jtj(w, f); ts.append((time.time() - t0) * 1e3)
    except Exception as e:
        print(mode, "FAILED:", e); continue
    res[mode] = (JTJ.copy(), JTf.copy())
    print("%-7s J^T J e2e ms: %s   finite=%s sym=%s" % (mode, ["%.2f" % t for t in ts], np.isfinite(JTJ).all(), np.max(np.abs(JTJ - JTJ.T)) == 0.0))
os.environ.pop("B200_JTJ", None)
if "dmma" in res:
    R, Rf = res["dmma"]
    sc = np.max(np.abs(R))
    for mode in ("ozaki", "ozaki7"):
        if mode not in res: continue
        X, Xf = res[mode]
        d = np.abs(X - R)
        print("%s vs dmma: max abs diff / max|R| = %.3e ; max rel elementwise (|R| > 1e-6 max) = %.3e ; jtf rel = %.3e" % (
            mode, d.max() / sc, np.max(d[np.abs(R) > 1e-6 * sc] / np.abs(R[np.abs(R) > 1e-6 * sc])), np.max(np.abs(Xf - Rf)) / np.max(np.abs(Rf))))
        if d.max() / sc > 1e-9:
            # error map per 64 x 64 block
            nb = (Np + 63) // 64
            m = np.zeros((nb, nb))
            for i in range(nb):
                for j in range(nb):
                    m[i, j] = d[i*64:(i+1)*64, j*64:(j+1)*64].max() / sc
            np.set_printoptions(linewidth=250, precision=1)
            print("block error map (rows i, cols j):"); print(m)
            print("sample R[0,:4]", R[0, :4], "X[0,:4]", X[0, :4]); print("sample R[200,:4]", R[200, :4], "X[200,:4]", X[200, :4])
    # accuracy against a long-double host reference on a few columns
    Jh = np.empty((nE, Np)); at.fill_dprobs(Jh, row_scale=w)
    cols = [0, 1, 17, 300, Np - 1]
    ref = (Jh[:, cols].astype(np.longdouble).T @ Jh.astype(np.longdouble)).astype(np.float64)
    for mode in res:
        print("%-7s vs long-double host rows %s: max abs / max|R| = %.3e" % (mode, cols, np.max(np.abs(res[mode][0][cols] - ref)) / sc))
