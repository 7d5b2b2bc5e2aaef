#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_c27.log
timeout 300 python tools/time_c3.py 2000 > gpurun_out/c3_2000_c27.log 2>&1
timeout 600 python tools/time_c3.py 50000 > gpurun_out/c3_50000_c27.log 2>&1
tail -n 3 gpurun_out/pytest_c27.log; grep "dprobs device\|max|J" gpurun_out/c3_2000_c27.log gpurun_out/c3_50000_c27.log
