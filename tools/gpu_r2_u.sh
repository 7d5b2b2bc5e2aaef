#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2u_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
tail -4 gpurun_out/r2u_pytest.log; tail -3 gpurun_out/r2u_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
x=d['extra_configs']['c3_d64_dprobs']; print({k:x[k] for k in ('ms','tflops','frac','hbm_frac','speedup_vs_dense_level_path')}); print(x['dense_level_path']['ms_fill_only'], x['parity'])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['jtj']['ms'])
PY
