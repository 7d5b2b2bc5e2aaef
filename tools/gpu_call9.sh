#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c9.log
B200_ACC_ST256=1 B200_ACC_RSUB=8 timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c9b.log
timeout 600 python tools/qt_sweep.py "" "B200_CHAIN_EARLY=0" "B200_CHAIN_K=3" "B200_CHAIN_K=4" "B200_ACC_ST256=1" "B200_ACC_RSUB=8" "B200_ACC_RSUB=8,B200_ACC_CHUNK=15" "B200_ACC_RSUB=8,B200_ACC_ST256=1" "B200_ACC_RSUB=16,B200_ACC_CHUNK=5" "B200_ACC_RSUB=4,B200_ACC_CHUNK=20" "B200_ACC_RSUB=8,B200_ACC_ST256=1,B200_DBG=2" "B200_ACC_RSUB=8,B200_ACC_ST256=1,B200_DBG=1" "B200_CHAIN_PROF=1" > gpurun_out/qt_c9.log 2>&1
tail -n 3 gpurun_out/pytest_c9.log gpurun_out/pytest_c9b.log; grep -v "prof\]" gpurun_out/qt_c9.log; grep "chain prof" gpurun_out/qt_c9.log | tail -2; grep "accum prof" gpurun_out/qt_c9.log | tail -1
