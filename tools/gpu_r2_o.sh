#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/oz_test.py c2_full_layout 3 > gpurun_out/r2o_oz.log 2>&1
tail -12 gpurun_out/r2o_oz.log
B200_JTJ=ozaki timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_launches_oz.csv python tools/oz_one.py > gpurun_out/r2o_oz_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2o_launches_oz.csv 2>/dev/null | tail -12
