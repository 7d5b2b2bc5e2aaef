#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2g_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3 50000 > gpurun_out/r2g_c3.log 2>&1
REPS=5 timeout 300 python tools/prof_r2.py c5 > gpurun_out/r2g_c5.log 2>&1
M="--metrics gpu__time_duration.sum --clock-control none --csv"
REPS=1 timeout 300 ncu $M -c 2000 --log-file gpurun_out/r2g_launches_c3.csv python tools/prof_r2.py c3 50000 > gpurun_out/r2g_c3n.log 2>&1
tail -4 gpurun_out/r2g_pytest.log; grep -v "^==" gpurun_out/r2g_c3.log gpurun_out/r2g_c5.log | tail
