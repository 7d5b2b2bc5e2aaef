#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2q_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -4 gpurun_out/r2q_pytest.log; tail -3 gpurun_out/r2q_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['jtj'],indent=1)[:2500]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
