#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pygsti_dropin.py tests/test_gpu_synthetic.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2x_pytest.log; tail -4 gpurun_out/r2x_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; tail -2 gpurun_out/r2x_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
x=d['extra_configs']; print('c5', x['c5_d256_probs']['ms'], x['c5_d256_probs']['embedded_model']); print('c4 lind', x['c4_cptplnd_hessian'].get('lindblad_members',{}).get('e2e_ms'), 'c3', x['c3_d64_dprobs']['ms'])
"
timeout 300 python - <<'PY' > gpurun_out/r2x_lind.log 2>&1
import sys, time, numpy as np
sys.path.insert(0, '.')
from types import SimpleNamespace as NS
from pygsti_b200 import engine
ctx = engine.Context(0)
rng = np.random.default_rng(11)
egs = []
for _ in range(7):
    B = (rng.standard_normal((240, 16, 16)) + 1j * rng.standard_normal((240, 16, 16))) / 16
    egs.append(NS(B_re=np.ascontiguousarray(B.real), B_im=np.ascontiguousarray(B.imag), c=0.02 * (rng.standard_normal(240) + 1j * rng.standard_normal(240)),
                  dc=rng.standard_normal((240, 240)) + 1j * rng.standard_normal((240, 240))))
mem = [NS(kind="op", errgen=g, static=rng.standard_normal((16, 16))) for g in range(5)] + [NS(kind="rho", errgen=5, static=rng.standard_normal(16))] + [NS(kind="eff", errgen=6, static=rng.standard_normal(16)) for _ in range(4)]
ctx.lindblad_members(16, egs, mem)
ts = []
for _ in range(3):
    t0 = time.time(); ctx.lindblad_members(16, egs, mem); ts.append((time.time() - t0) * 1e3)
print("lindblad_members e2e ms", ts)
PY
tail -2 gpurun_out/r2x_lind.log
