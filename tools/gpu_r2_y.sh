#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -3 gpurun_out/r2y_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('e2e_plugin'), indent=1)[:1800]); print(d['value'], d['jtj']['ms'])
"
