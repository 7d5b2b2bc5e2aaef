#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c13.log
timeout 600 python tools/qt_sweep.py "B200_VERBOSE=1" "B200_L2_PERSIST=0" "B200_L2_PERSIST_MB=32" "B200_L2_PERSIST_MB=64" "B200_ACC_RSUB=8" "B200_ACC_RSUB=8,B200_L2_PERSIST=0" "B200_ACC_ST256=1" "B200_ACC_ST256=1,B200_ACC_RSUB=8" "B200_UNIT_OUTCOMES=2" "B200_UNIT_OUTCOMES=2,B200_ACC_RSUB=8" "B200_CHAIN_CTAS=3" "B200_DBG=2" "B200_DBG=1" > gpurun_out/qt_c13.log 2>&1
tail -n 3 gpurun_out/pytest_c13.log; cat gpurun_out/qt_c13.log
