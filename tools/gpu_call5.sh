#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_c5.log
( python tools/qt_dev.py
  B200_ACC_V6=1 python tools/qt_dev.py
  B200_UNIT_OUTCOMES=2 python tools/qt_dev.py
  B200_CHAIN_PROF=1 python tools/qt_dev.py ) > gpurun_out/qt_v7.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_c3_50000.csv python tools/time_c3.py 50000 > gpurun_out/c3_50000_ncu.log 2>&1
tail -n 12 gpurun_out/pytest_c5.log; cat gpurun_out/qt_v7.log
