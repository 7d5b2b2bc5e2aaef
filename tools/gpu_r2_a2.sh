#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2a2_pytest.log; tail -2 gpurun_out/r2a2_pytest.log
timeout 300 python tools/time_c5.py > gpurun_out/r2a2_c5.log 2>&1; tail -3 gpurun_out/r2a2_c5.log
B200_LEVEL_LAUNCHES=1 timeout 300 python tools/time_c5.py > gpurun_out/r2a2_c5_launches.log 2>&1; tail -3 gpurun_out/r2a2_c5_launches.log
