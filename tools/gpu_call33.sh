#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_c33.log
timeout 300 python tools/time_c3.py 2000 > gpurun_out/c3_2000_c33.log 2>&1
timeout 600 python tools/time_c3.py 50000 > gpurun_out/c3_50000_c33.log 2>&1
tail -n 2 gpurun_out/pytest_c33.log; grep "dprobs device\|max|J" gpurun_out/c3_2000_c33.log gpurun_out/c3_50000_c33.log
