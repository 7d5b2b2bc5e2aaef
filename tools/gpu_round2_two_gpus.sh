#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2v_pytest_multi.log
tail -3 gpurun_out/r2v_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err
tail -3 gpurun_out/r2v_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench_n2.json').read().strip().splitlines()[-1])
x=d['extra_configs']['c3_d64_dprobs']; print({k:x[k] for k in ('ms','ms_fill_only','parallelism')})
print(d['value'], d['ms_per_step'], d['jtj']['ms'], d['jtj']['ms_local'], d['e2e']['ms_per_step'], d['multi_gpu']['fill_only'], d['multi_gpu']['allgather']['ms'])
PY
