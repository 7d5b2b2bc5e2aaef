"""Timing of the generic (d = 64 / 256) paths on BASELINE-like samples (dev tool).
C3 sample: the 3-qubit LocalNoiseModel of tests/golden/c3_3q_localnoise_sub.npz (real G, real sparse D, Np = 775) with
2000 random circuits of depth U{1..256} (reference 480 ms probs, ~357 s dprobs on 1 core).
C5 sample: d = 256 random dense model, 14 layer labels, 500 random circuits depth U{1..64}, 16 outcomes (reference 177 ms)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case
from tests import synth
from oracle import oracle_c

ctx = engine.Context(0)
orc = oracle_c.Oracle("port")
rng = np.random.default_rng(0)

def timeit(f, n=3):
    f(); ctx.sync(); ts = []
    for _ in range(n):
        t0 = time.time(); f(); ctx.sync(); ts.append(time.time() - t0)
    return min(ts)

# ---- C3 sample ----
c = Case("c3_3q_localnoise_sub"); a = c.atoms[0]
n_ops, n_eff = a["tables"].n_ops, a["tables"].n_eff
ncirc = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
circs = [(0, [int(x) for x in rng.integers(0, n_ops, size=int(rng.integers(1, 257)))], list(range(n_eff))) for _ in range(ncirc)]
t = synth.make_tables(64, n_ops, 1, n_eff, circs, use_cache=False)
at = ctx.upload_atom(t); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
print("C3 sample:", at.info(), "D nnz", a["D"].rows.size)
p = np.empty(t.n_elements)
tp = timeit(lambda: at.fill_probs(p))
print("  GPU probs  %.2f ms  (%.3e outcomes/s)" % (tp * 1e3, t.n_elements / tp))
J = engine.pinned_empty((t.n_elements, a["D"].n_params))
td = timeit(lambda: at.fill_dprobs(J), n=2)
print("  GPU dprobs %.1f ms  (%.3e outcomes/s, %.3e dprobs-el/s)" % (td * 1e3, t.n_elements / td, t.n_elements * a["D"].n_params / td))
t0 = time.time(); po = orc.mapfill_probs(t, a["G"], a["rho"], a["E"]); tc = time.time() - t0
print("  CPU oracle probs %.1f ms (1 core); max|p-po| = %.2e" % (tc * 1e3, np.max(np.abs(p - po))))
sub = synth.make_tables(64, n_ops, 1, n_eff, circs[:20], use_cache=False)
Jo, _ = orc.dprobs_analytic(sub, a["G"], a["rho"], a["E"], a["D"])
print("  max|J - oracle| on first 20 circuits = %.2e" % np.max(np.abs(J[:sub.n_elements] - Jo)))
at.free()

# ---- C5 sample ----
G, rho, E = synth.random_model(256, 14, 1, 16, seed=1)
circs = [(0, [int(x) for x in rng.integers(0, 14, size=int(rng.integers(1, 65)))], list(range(16))) for _ in range(500)]
t = synth.make_tables(256, 14, 1, 16, circs, use_cache=False)
at = ctx.upload_atom(t); at.set_model(G, rho, E)
p = np.empty(t.n_elements)
tp = timeit(lambda: at.fill_probs(p))
print("C5 sample:", at.info())
print("  GPU probs  %.2f ms  (%.3e outcomes/s)" % (tp * 1e3, t.n_elements / tp))
t0 = time.time(); po = orc.mapfill_probs(t, G, rho, E); tc = time.time() - t0
print("  CPU oracle probs %.1f ms (1 core); max|p-po| = %.2e" % (tc * 1e3, np.max(np.abs(p - po))))

# ---- C5 at BASELINE size: 5000 circuits, depth U{1..128}: FP64 tensor (DMMA) throughput of the level-batched path ----
import torch
circs = [(0, [int(x) for x in rng.integers(0, 14, size=int(rng.integers(1, 129)))], list(range(16))) for _ in range(5000)]
t = synth.make_tables(256, 14, 1, 16, circs, use_cache=False)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx2 = engine.Context(0, stream=stream.cuda_stream)
at = ctx2.upload_atom(t); at.set_model(G, rho, E)
P = torch.empty(t.n_elements, dtype=torch.float64, device="cuda")
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream); at.fill_probs_dev(P.data_ptr()); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
info = at.info(); flops = 2.0 * 256 * 256 * info["n_prop_expanded"]
print("C5 full size:", info)
print("  GPU probs (device) %.2f ms -> %.2f TFLOP/s FP64 (%.3e outcomes/s); launches/level ~%d levels" % (min(ts), flops / (min(ts) * 1e-3) / 1e12, t.n_elements / (min(ts) * 1e-3), info["max_depth"]))
po = orc.mapfill_probs(synth.make_tables(256, 14, 1, 16, circs[:40], use_cache=False), G, rho, E)
print("  max|p - oracle| (first 40 circuits) = %.2e" % np.max(np.abs(P[:po.size].cpu().numpy() - po)))
