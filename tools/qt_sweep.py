"""Device-side timing of probs / dprobs on the bench layout for several env-knob settings in ONE process (dev tool).
usage: qt_sweep.py "K1=V1,K2=V2" "K3=V3" ...   ("" = defaults); the engine reads its B200_* knobs on every call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case
c = Case("c2_full_layout"); a = c.atoms[0]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream=stream.cuda_stream)
nE, Np = c.n_elements, c.num_params
J = torch.empty((nE, Np), dtype=torch.float64, device="cuda"); p = torch.empty(nE, dtype=torch.float64, device="cuda")
rows = torch.as_tensor(c["dprobs_matrix_sample_elements"], device="cuda")
for spec in (sys.argv[1:] or [""]):
    kv = dict(x.split("=") for x in spec.split(",") if x)
    for k in [k for k in os.environ if k.startswith("B200_")]: del os.environ[k]
    os.environ.update(kv)
    at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])   # (upload-time knobs too)
    res = {}
    for what in ("probs", "dprobs"):
        ts = []
        J.zero_()
        for r in range(14):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if what == "probs": at.fill_probs_dev(p.data_ptr())
            else: at.fill_dprobs_dev(J.data_ptr(), Np, p.data_ptr())
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[what] = (min(ts[3:]), float(np.median(ts[3:])))
    err = float(np.max(np.abs(J[rows].cpu().numpy() - c["dprobs_matrix_sample_rows"])))
    print("%-40s probs %.3f/%.3f ms  dprobs %.3f/%.3f ms (min/median) -> %.0f GB/s  err %.1e" %
          (spec or "(defaults)", res["probs"][0], res["probs"][1], res["dprobs"][0], res["dprobs"][1], nE * (Np + 1) * 8 / res["dprobs"][0] / 1e6, err), flush=True)
