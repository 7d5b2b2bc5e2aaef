#!/bin/bash
# round 2, call A (one B200): full GPU test suite (not -x: see everything that fails), the bench with all extra configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -n 25 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
