#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fj_backward -c 1 -o gpurun_out/r2s_fjb python tools/prof_r2.py c3f 10000 > gpurun_out/r2s_ncu.log 2>&1
tail -3 gpurun_out/r2s_ncu.log
