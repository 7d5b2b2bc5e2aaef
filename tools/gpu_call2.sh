#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_all.log
( python tools/qt_dev.py
  B200_CHAIN_K=1 python tools/qt_dev.py
  B200_CHAIN_K=4 python tools/qt_dev.py
  B200_CHAIN_K=8 python tools/qt_dev.py
  B200_CHAIN_CTAS=6 python tools/qt_dev.py
  B200_CHAIN_CTAS=6 B200_CHAIN_K=4 python tools/qt_dev.py
  B200_ACC_ST256=0 python tools/qt_dev.py
  B200_UNIT_OUTCOMES=2 python tools/qt_dev.py ) > gpurun_out/qt_matrix.log 2>&1
tail -n 30 gpurun_out/pytest_all.log; cat gpurun_out/qt_matrix.log
