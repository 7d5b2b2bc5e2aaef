#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2d_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c4 > gpurun_out/r2d_c4.log 2>&1
REPS=3 timeout 300 python tools/prof_r2.py jtj > gpurun_out/r2d_jtj.log 2>&1
tail -5 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_c4.log gpurun_out/r2d_jtj.log
