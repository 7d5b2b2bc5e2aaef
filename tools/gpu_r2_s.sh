#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "factored" 2>&1 | tail -25 > gpurun_out/r2s_pytest.log
tail -5 gpurun_out/r2s_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fj_ -c 2 -o gpurun_out/r2s_fj python tools/prof_r2.py c3f 10000 > gpurun_out/r2s_ncu.log 2>&1
tail -3 gpurun_out/r2s_ncu.log
