#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c17.log
timeout 600 python tools/qt_sweep.py "" "B200_ACC_PFD=0" "B200_ACC_PFD=40" "B200_ACC_PFD=400" "B200_ACC_PFD=1000" "B200_ACC_RSUB=8" "B200_ACC_ST256=1" "B200_UNIT_OUTCOMES=2" "B200_DBG=1" "B200_DBG=2" "B200_CHAIN_PROF=1" > gpurun_out/qt_c17.log 2>&1
tail -n 3 gpurun_out/pytest_c17.log; grep -v "prof\]" gpurun_out/qt_c17.log; grep "chain prof" gpurun_out/qt_c17.log | tail -2; grep "accum prof" gpurun_out/qt_c17.log | tail -1
