#!/bin/bash
mkdir -p gpurun_out
B200_ACC_TMA=1 timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c26.log
timeout 600 python tools/qt_sweep.py "" "B200_ACC_TMA=1" "B200_ACC_TMA=1,B200_ACC_RSUB=8" "B200_ACC_TMA=1,B200_DBG=1" "B200_ACC_TMA=1,B200_ACC_CHUNK=32" "B200_ACC_TMA=1" "" > gpurun_out/qt_c26.log 2>&1
tail -n 3 gpurun_out/pytest_c26.log; cat gpurun_out/qt_c26.log
