"""BASELINE config 5 on one GPU (dev tool): d = 256 dense model (14 layer labels, 16 outcomes), 5000 random circuits of depth
U{1..128}, probs only -- the level-batched FP64 tensor-core (DMMA) path; device-resident time with CUDA events, FP64 TFLOP/s
against the measured DMMA peak (tools/ubench_fp64.cu: 37.2 TFLOP/s), parity of the first circuits against the C oracle."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygsti_b200 import engine
from tests import synth
from oracle import oracle_c

n_circ = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(0)
G, rho, E = synth.random_model(256, 14, 1, 16, seed=1)
circs = [(0, [int(x) for x in rng.integers(0, 14, size=int(rng.integers(1, 129)))], list(range(16))) for _ in range(n_circ)]
t = synth.make_tables(256, 14, 1, 16, circs, use_cache=False)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = engine.Context(0, stream=stream.cuda_stream)
at = ctx.upload_atom(t); at.set_model(G, rho, E)
P = torch.empty(t.n_elements, dtype=torch.float64, device="cuda")
ts = []
for _ in range(6):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream); at.fill_probs_dev(P.data_ptr()); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
info = at.info(); flops = 2.0 * 256 * 256 * info["n_prop_expanded"]
ms = min(ts[1:])
print("C5 x %d circuits: %s" % (n_circ, info))
print("  probs device-resident %.2f ms -> %.2f TFLOP/s FP64 = %.2f of the measured DMMA peak (37.2); %.3e outcomes/s; %d level launches"
      % (ms, flops / (ms * 1e-3) / 1e12, flops / (ms * 1e-3) / 1e12 / 37.2, t.n_elements / (ms * 1e-3), info["max_depth"]))
orc = oracle_c.Oracle("port")
po = orc.mapfill_probs(synth.make_tables(256, 14, 1, 16, circs[:40], use_cache=False), G, rho, E)
print("  max|p - oracle| (first 40 circuits) = %.2e" % np.max(np.abs(P[:po.size].cpu().numpy() - po)))
