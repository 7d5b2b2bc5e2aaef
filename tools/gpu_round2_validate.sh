#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z2_smoke.log 2>&1; tail -1 gpurun_out/r2z2_smoke.log | cut -c1-300
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2z2_pytest.log; tail -2 gpurun_out/r2z2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err; tail -2 gpurun_out/r2z2_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2z2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['jtj']['ms'], d['e2e']['ms_per_step'], d['e2e_plugin'].get('ratio'))
x=d['extra_configs']; print(x['c3_d64_dprobs']['ms'], x['c4_cptplnd_hessian']['hessian_rectangle']['e2e_ms'], x['c4_cptplnd_hessian']['lindblad_members'].get('e2e_ms'), x['c5_d256_probs']['ms'], x['c5_d256_probs']['embedded_model'].get('ms_factored'))
"
