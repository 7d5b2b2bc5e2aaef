"""Device-side timing of probs / dprobs on the bench layout for the current env knobs (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case
c = Case(sys.argv[1] if len(sys.argv) > 1 else "c2_full_layout"); a = c.atoms[0]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream=stream.cuda_stream)
at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
nE, Np = c.n_elements, c.num_params
J = torch.empty((nE, Np), dtype=torch.float64, device="cuda"); p = torch.empty(nE, dtype=torch.float64, device="cuda")
res = {}
for what in ("probs", "dprobs"):
    ts = []
    for r in range(12):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if what == "probs": at.fill_probs_dev(p.data_ptr())
        else: at.fill_dprobs_dev(J.data_ptr(), Np, p.data_ptr())
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[what] = (min(ts[2:]), float(np.median(ts[2:])))
rows = torch.as_tensor(c["dprobs_matrix_sample_elements"], device="cuda")
err = float(np.max(np.abs(J[rows].cpu().numpy() - c["dprobs_matrix_sample_rows"])))
knobs = {k: v for k, v in os.environ.items() if k.startswith("B200_")}
print("%s probs %.3f/%.3f ms  dprobs %.3f/%.3f ms (min/median) -> %.0f GB/s  err %.1e" %
      (knobs, res["probs"][0], res["probs"][1], res["dprobs"][0], res["dprobs"][1], nE * (Np + 1) * 8 / res["dprobs"][0] / 1e6, err))
