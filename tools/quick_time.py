"""Quick device-side timing of the probs / dprobs kernels on a golden layout (dev tool, not the bench)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case

name = sys.argv[1] if len(sys.argv) > 1 else "c2_full_layout"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
c = Case(name)
a = c.atoms[0]
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream=stream.cuda_stream)
at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
print(at.info())
nE, Np = c.n_elements, c.num_params
J = torch.empty((nE, Np), dtype=torch.float64, device="cuda")
p = torch.empty(nE, dtype=torch.float64, device="cuda")
for what in ("probs", "dprobs"):
    ts = []
    for r in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if what == "probs": at.fill_probs_dev(p.data_ptr())
        else: at.fill_dprobs_dev(J.data_ptr(), Np, p.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[2:])
    print("%s: %.3f ms (all: %s)  outcomes/s=%.3e" % (what, t, ["%.3f" % x for x in ts], nE / (t * 1e-3)))
    if what == "dprobs":
        print("   bytes=%.3f GB -> %.1f GB/s" % (nE * (Np + 1) * 8 / 1e9, nE * (Np + 1) * 8 / (t * 1e-3) / 1e9))
# host-buffer e2e
Jh = engine.pinned_empty((nE, Np)); ph = np.empty(nE)
for r in range(3):
    t0 = time.time(); at.fill_dprobs(Jh, ph); t1 = time.time()
    print("e2e pinned host dprobs: %.1f ms" % ((t1 - t0) * 1e3))
Jp = np.empty((nE, Np))
for r in range(2):
    t0 = time.time(); at.fill_dprobs(Jp, ph); t1 = time.time()
    print("e2e pageable host dprobs: %.1f ms" % ((t1 - t0) * 1e3))
# fused objective pieces (8f-1): scaled Jacobian into pinned host memory, and J^T J / J^T f only
w = np.random.default_rng(0).uniform(0.5, 1.5, nE); f = np.random.default_rng(1).standard_normal(nE)
for r in range(2):
    t0 = time.time(); at.fill_dprobs(Jh, row_scale=w); t1 = time.time()
    print("e2e scaled dprobs (pinned): %.1f ms" % ((t1 - t0) * 1e3))
for r in range(3):
    t0 = time.time(); JTJ, JTf = at.jtj(w, f); t1 = time.time()
    print("e2e J^T J + J^T f (Jacobian stays on the device): %.1f ms" % ((t1 - t0) * 1e3))
chk = Jh[:, :5].T @ Jh
print("   max |JTJ[:5] - host| rel = %.2e" % (np.max(np.abs(JTJ[:5] - chk)) / np.max(np.abs(chk))))
