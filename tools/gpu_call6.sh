#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_c6.log
python tools/qt_dev.py > gpurun_out/qt_c6.log 2>&1
timeout 600 python tools/time_c3.py 2000 > gpurun_out/c3_2000_v3.log 2>&1
timeout 900 python tools/time_c3.py 50000 > gpurun_out/c3_50000_v3.log 2>&1
tail -n 8 gpurun_out/pytest_c6.log; cat gpurun_out/qt_c6.log gpurun_out/c3_2000_v3.log gpurun_out/c3_50000_v3.log
