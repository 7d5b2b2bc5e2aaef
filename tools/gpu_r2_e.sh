#!/bin/bash
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2e_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3 2000 > gpurun_out/r2e_c3.log 2>&1
REPS=3 timeout 300 python tools/prof_r2.py c3 50000 >> gpurun_out/r2e_c3.log 2>&1
REPS=2 timeout 300 ncu $M -c 600 --log-file gpurun_out/r2e_launches_c4.csv python tools/prof_r2.py c4 > gpurun_out/r2e_c4.log 2>&1
REPS=2 timeout 300 ncu $M -c 100 --log-file gpurun_out/r2e_launches_jtj.csv python tools/prof_r2.py jtj > gpurun_out/r2e_jtj.log 2>&1
REPS=1 timeout 300 ncu $M -c 2000 --log-file gpurun_out/r2e_launches_c3.csv python tools/prof_r2.py c3 50000 > gpurun_out/r2e_c3n.log 2>&1
tail -5 gpurun_out/r2e_pytest.log; grep -v "^==" gpurun_out/r2e_c3.log gpurun_out/r2e_c4.log gpurun_out/r2e_jtj.log | tail
