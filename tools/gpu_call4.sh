#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_c4.log
B200_CHAIN_PROF=1 timeout 300 python tools/qt_dev.py 2>&1 | tail -4 > gpurun_out/accum_prof.log
timeout 600 python tools/time_c3.py 2000 > gpurun_out/c3_2000_v2.log 2>&1
timeout 900 python tools/time_c3.py 50000 > gpurun_out/c3_50000_v2.log 2>&1
B200_LJ_ACCUM_V1=1 timeout 600 python tools/time_c3.py 2000 > gpurun_out/c3_2000_v1.log 2>&1
tail -n 12 gpurun_out/pytest_c4.log; cat gpurun_out/accum_prof.log gpurun_out/c3_2000_v2.log gpurun_out/c3_50000_v2.log gpurun_out/c3_2000_v1.log
