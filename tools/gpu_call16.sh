#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accum_trie -s 4 -c 1 -f -o gpurun_out/accum_r01c python tools/qt_sweep.py "" > gpurun_out/ncu_accum_r01c.log 2>&1
timeout 300 python tools/qt_sweep.py "B200_ACC_ST256=1" "B200_ACC_ST256=1,B200_ACC_RSUB=8" "B200_ACC_RSUB=8" "" >> gpurun_out/ncu_accum_r01c.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -5 gpurun_out/ncu_accum_r01c.log
