#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pygsti_dropin.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_c3.log
B200_CHAIN_PROF=1 timeout 300 python tools/qt_dev.py > gpurun_out/chain_prof.log 2>&1
timeout 300 python tools/qt_dev.py > gpurun_out/qt_default.log 2>&1
timeout 600 python tools/time_c3.py 2000 > gpurun_out/c3_2000.log 2>&1
timeout 900 python tools/time_c3.py 50000 > gpurun_out/c3_50000.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_c3_2000.csv python tools/time_c3.py 2000 > gpurun_out/c3_ncu.log 2>&1
tail -n 12 gpurun_out/pytest_c3.log; tail -n 12 gpurun_out/chain_prof.log; cat gpurun_out/qt_default.log gpurun_out/c3_2000.log gpurun_out/c3_50000.log
