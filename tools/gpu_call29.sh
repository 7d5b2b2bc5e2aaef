#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_c29.log
B200_UNIT_OUTCOMES=2 timeout 600 python -m pytest tests/test_gpu_synthetic.py -m gpu -q -x 2>&1 | tail -2 >> gpurun_out/pytest_c29.log
timeout 600 python tools/qt_sweep.py "" "B200_DBG=1" "B200_ACC_TMA=1" "" > gpurun_out/qt_c29.log 2>&1
tail -n 5 gpurun_out/pytest_c29.log; cat gpurun_out/qt_c29.log
