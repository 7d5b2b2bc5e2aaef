#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c18.log
B200_ACC_IDXW=0 timeout 600 python -m pytest tests/test_gpu_synthetic.py -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/pytest_c18.log
timeout 600 python tools/qt_sweep.py "" "B200_ACC_IDXW=0" "B200_ACC_RSUB=8" "B200_UNIT_OUTCOMES=2" "B200_DBG=1" "B200_DBG=2" "B200_ACC_IDXW=0,B200_DBG=2" "B200_CHAIN_PROF=1" "" > gpurun_out/qt_c18.log 2>&1
tail -n 6 gpurun_out/pytest_c18.log; grep -v "prof\]" gpurun_out/qt_c18.log; grep "accum prof" gpurun_out/qt_c18.log | tail -1
