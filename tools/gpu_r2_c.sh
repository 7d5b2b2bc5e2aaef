#!/bin/bash
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 ncu $M -c 400 --log-file gpurun_out/r2c_launches_c4.csv python tools/prof_r2.py c4 > gpurun_out/r2c_c4.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_atb_dmma -s 1 -c 1 -f -o gpurun_out/r2c_atb python tools/prof_r2.py jtj > gpurun_out/r2c_jtj.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_level_accum2 -s 1 -c 1 -f -o gpurun_out/r2c_accum2 python tools/prof_r2.py c3 2000 > gpurun_out/r2c_c3.log 2>&1
cat gpurun_out/r2c_c4.log gpurun_out/r2c_jtj.log gpurun_out/r2c_c3.log | grep -v "^==PROF" | tail -12
