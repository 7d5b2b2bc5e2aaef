"""BASELINE config 3 on one GPU (dev tool): 3-qubit crosstalk-free `full TP` model (d = 64, Np = 775; tensors from the
reference, tests/golden/c3_3q_localnoise_sub.npz), N random circuits of depth U{1..256} (rng seed 0, layers
drawn uniformly from the 10 primitive layer labels), bulk_fill_dprobs + probs.  Device-resident time with CUDA events,
end-to-end time into pinned host memory, and a parity check of a few circuits against the C oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case
from tests import synth
from oracle import oracle_c

n_circ = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
c = Case("c3_3q_localnoise_sub"); a = c.atoms[0]
n_ops, n_eff = a["tables"].n_ops, a["tables"].n_eff
rng = np.random.default_rng(0)
circs = [(0, [int(x) for x in rng.integers(0, n_ops, size=int(rng.integers(1, 257)))], list(range(n_eff))) for _ in range(n_circ)]
t0 = time.time(); t = synth.make_tables(64, n_ops, 1, n_eff, circs, use_cache=False); t_tab = time.time() - t0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream=stream.cuda_stream)
t0 = time.time(); at = ctx.upload_atom(t); t_up = time.time() - t0
at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
info = at.info(); nE, Np = t.n_elements, a["D"].n_params
print("C3 x %d circuits: %s  (tables %.1fs, upload %.1fs)" % (n_circ, info, t_tab, t_up))
J = torch.empty((nE, Np), dtype=torch.float64, device="cuda"); p = torch.empty(nE, dtype=torch.float64, device="cuda")
ts = []
for r in range(6):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); at.fill_dprobs_dev(J.data_ptr(), Np, p.data_ptr()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
td = min(ts[1:])
fl = 2.0 * 64 * 64 * info["n_prop_expanded"] * (1 + n_eff)
print("  dprobs device-resident: %.2f ms (all %s) -> %.3e outcomes/s, %.3e dprobs-el/s; sweeps %.1f GFLOP -> %.1f TFLOP/s if they were all of it; J %.2f GB"
      % (td, ["%.1f" % x for x in ts], nE / td * 1e3, nE * Np / td * 1e3, fl / 1e9, fl / td / 1e9, nE * Np * 8 / 1e9))
ts = []
for r in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); at.fill_probs_dev(p.data_ptr()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("  probs device-resident: %.2f ms -> %.3e outcomes/s" % (min(ts[1:]), nE / min(ts[1:]) * 1e3))
if nE * Np * 8 < 8e9:
    Jh = engine.pinned_empty((nE, Np)); ph = np.empty(nE)
    for r in range(3):
        t0 = time.time(); at.fill_dprobs(Jh, ph); t1 = time.time()
    print("  dprobs end to end (pinned host): %.1f ms -> %.3e outcomes/s" % ((t1 - t0) * 1e3, nE / (t1 - t0)))
orc = oracle_c.Oracle("port")
sub = synth.make_tables(64, n_ops, 1, n_eff, circs[:12], use_cache=False)
Jo, po = orc.dprobs_analytic(sub, a["G"], a["rho"], a["E"], a["D"])
Jg = J[:sub.n_elements].cpu().numpy()
print("  max|J - oracle| (first 12 circuits) = %.2e (scale %.2e); max|p - oracle| = %.2e" %
      (np.max(np.abs(Jg - Jo)), np.max(np.abs(Jo)), np.max(np.abs(p[:sub.n_elements].cpu().numpy() - po))))
