"""Dev tool (round 2): one pass over the secondary hot paths so that ncu can list / capture their kernels.
  python tools/prof_r2.py jtj   -> C2 full layout: J^T J (k_atb_dmma)
  python tools/prof_r2.py c4    -> C4 (CPTPLND, maxL=16): dprobs + one 64x64 Hessian rectangle
  python tools/prof_r2.py c3 N  -> C3 with N circuits: dprobs
  python tools/prof_r2.py c5    -> C5: probs"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygsti_b200 import engine, fixtures as fx

what = sys.argv[1] if len(sys.argv) > 1 else "jtj"
reps = int(os.environ.get("REPS", "2"))
ctx = engine.Context(0)

def wall(fn, n=reps):
    fn(); ctx.sync(); ts = []
    for _ in range(n):
        t0 = time.time(); fn(); ctx.sync(); ts.append(time.time() - t0)
    return min(ts) * 1e3

if what == "jtj":
    c = fx.Case("c2_full_layout"); a = c.atoms[0]
    at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
    rs = np.random.default_rng(0).uniform(0.5, 1.5, c.n_elements); f = np.random.default_rng(1).standard_normal(c.n_elements)
    print("jtj ms (incl. 15 MB D2H):", wall(lambda: at.jtj(rs, f)))
elif what == "c4":
    c = fx.Case("c4_gst16_layout"); a = c.atoms[0]
    at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
    nE, Np = c.n_elements, c.num_params
    J = engine.pinned_empty((nE, Np))
    print("c4 dprobs e2e ms:", wall(lambda: at.fill_dprobs(J)))
    r = c["hess_rects"][0]; p1, p2 = np.arange(r[0], r[1]), np.arange(r[2], r[3])
    w_h, w_d = np.random.default_rng(0).standard_normal(nE), np.random.default_rng(1).uniform(0.5, 1.5, nE)
    H2 = c.hess_map("H2r0")
    print("c4 hessian block ms:", wall(lambda: at.hessian_block(p1, p2, w_h, w_d, H2)), "info", at.info())
elif what == "c3":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    c = fx.Case("c3_3q_localnoise_sub"); a = c.atoms[0]
    t, _ = fx.random_layout(64, a["tables"].n_ops, a["tables"].n_eff, 50000, 256, seed=0, rows=(0, n))
    at = ctx.upload_atom(t); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
    import torch
    J = torch.empty((t.n_elements, a["D"].n_params), dtype=torch.float64, device="cuda")
    print("c3 x%d dprobs ms:" % n, wall(lambda: at.fill_dprobs_dev(J.data_ptr(), a["D"].n_params)))
elif what == "c3f":            # config 3 through the factor programs (kernels_factoredj.cuh) vs the dense level-batched path
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    c = fx.Case("c3_3q_localnoise_sub"); a = c.atoms[0]
    t, _ = fx.random_layout(64, a["tables"].n_ops, a["tables"].n_eff, 50000, 256, seed=0, rows=(0, n))
    Np = a["D"].n_params
    import torch
    at = ctx.upload_atom(t); at.set_model_factored(a["fm"]); at.set_derivs(a["D"])
    Jd = torch.empty((t.n_elements, Np), dtype=torch.float64, device="cuda"); Pd = torch.empty(t.n_elements, dtype=torch.float64, device="cuda")
    print("c3 x%d dense level path ms:" % n, wall(lambda: at.fill_dprobs_dev(Jd.data_ptr(), Np, Pd.data_ptr())))
    at.set_derivs_factored(a["Df"])
    Jf = torch.empty_like(Jd); Pf = torch.empty_like(Pd)
    print("c3 x%d factored path ms:" % n, wall(lambda: at.fill_dprobs_dev(Jf.data_ptr(), Np, Pf.data_ptr())))
    print("   max |J_f - J_dense| = %.3e (max |J| %.3e), max |p_f - p_dense| = %.3e" % (float((Jf - Jd).abs().max()), float(Jd.abs().max()), float((Pf - Pd).abs().max())))
elif what == "c5":
    G, rho, E = fx.random_dense_model(256, 14, 1, 16, seed=1)
    t, _ = fx.random_layout(256, 14, 16, 5000, 128, seed=0)
    at = ctx.upload_atom(t); at.set_model(G, rho, E)
    import torch
    P = torch.empty(t.n_elements, dtype=torch.float64, device="cuda")
    print("c5 probs ms:", wall(lambda: at.fill_probs_dev(P.data_ptr())))
