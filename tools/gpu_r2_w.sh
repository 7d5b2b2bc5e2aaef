#!/bin/bash
mkdir -p gpurun_out
B200_FJ_SINGLE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fj64_backward -c 1 -o gpurun_out/r2w_fj64_single python tools/prof_r2.py c3f 10000 > gpurun_out/r2w_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fj64_backward2 -c 1 -o gpurun_out/r2w_fj64_pair python tools/prof_r2.py c3f 10000 > gpurun_out/r2w_ncu2.log 2>&1
tail -2 gpurun_out/r2w_ncu1.log gpurun_out/r2w_ncu2.log
