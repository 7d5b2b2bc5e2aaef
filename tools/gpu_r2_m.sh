#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2m_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3 50000 > gpurun_out/r2m_c3.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -4 gpurun_out/r2m_pytest.log; grep -v "^==" gpurun_out/r2m_c3.log | tail -2; tail -2 gpurun_out/r2m_bench.err
