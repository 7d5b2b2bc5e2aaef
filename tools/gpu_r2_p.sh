#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_oz.csv python tools/oz_test.py c2_full_layout 1 > gpurun_out/r2p_oz_ncu.log 2>&1
B200_JTJ=ozaki ncu --set full --clock-control none --import-source on -k regex:k_oz_ -c 5 -o gpurun_out/r2p_oz python tools/oz_one.py > gpurun_out/r2p_oz_full.log 2>&1
python tools/launch_summary.py gpurun_out/r2p_launches_oz.csv 2>/dev/null | tail -30
