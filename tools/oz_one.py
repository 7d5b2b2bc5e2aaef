"""One J^T J call on the full config-2 layout (for ncu captures; mode from B200_JTJ)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygsti_b200 import engine
from pygsti_b200.fixtures import Case
c = Case(sys.argv[1] if len(sys.argv) > 1 else "c2_full_layout")
a = c.atoms[0]
ctx = engine.Context(0)
at = ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
rng = np.random.default_rng(0)
w = rng.uniform(0.5, 1.5, c.n_elements); f = rng.standard_normal(c.n_elements)
JTJ, JTf = at.jtj(w, f)
print(np.isfinite(JTJ).all())
