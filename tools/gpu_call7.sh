#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_level_accum2 -s 1 -c 1 -f -o gpurun_out/accum2_c3_2000 python tools/time_c3.py 2000 > gpurun_out/ncu_accum2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_accum2.log
