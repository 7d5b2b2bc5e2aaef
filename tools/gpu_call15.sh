#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/qt_sweep.py "" "B200_DBG=3" "B200_DBG=4" "B200_DBG=5" "B200_DBG=2" "B200_DBG=1" "B200_DBG=5,B200_ACC_ST256=1" "B200_DBG=3,B200_UNIT_OUTCOMES=2" "B200_UNIT_OUTCOMES=2" > gpurun_out/qt_c15.log 2>&1
cat gpurun_out/qt_c15.log
