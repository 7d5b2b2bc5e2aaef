#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c21.log
timeout 300 python tools/time_c3.py 2000 > gpurun_out/c3_2000_c21.log 2>&1
B200_LJ_PT=256 timeout 300 python tools/time_c3.py 2000 > gpurun_out/c3_2000_c21_pt256.log 2>&1
timeout 600 python tools/time_c3.py 50000 > gpurun_out/c3_50000_c21.log 2>&1
timeout 600 python tools/time_big_d.py 500 > gpurun_out/bigd_c21.log 2>&1
tail -n 3 gpurun_out/pytest_c21.log; grep "dprobs device\|probs device\|max|J" gpurun_out/c3_2000_c21.log gpurun_out/c3_2000_c21_pt256.log gpurun_out/c3_50000_c21.log; grep "C5\|GPU probs\|max|p" gpurun_out/bigd_c21.log
