#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2j_pytest.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_level_accum3 -s 1 -c 1 -f -o gpurun_out/r2j_accum3 python tools/prof_r2.py c3 4000 > gpurun_out/r2j_c3.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -5 gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print(d["ms_per_step"], d["extra_configs"]["c5_d256_probs"].get("embedded_model"), d["extra_configs"]["c5_d256_probs"]["ms"], d["extra_configs"]["c3_d64_dprobs"]["ms"])
PY
