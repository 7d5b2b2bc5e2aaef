#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "factored or config3" 2>&1 | tail -25 > gpurun_out/r2t_pytest.log
tail -3 gpurun_out/r2t_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3f 50000 > gpurun_out/r2t_c3f.log 2>&1; tail -2 gpurun_out/r2t_c3f.log
B200_FJ_SINGLE=1 REPS=3 timeout 300 python tools/prof_r2.py c3f 50000 > gpurun_out/r2t_c3f_single.log 2>&1; tail -2 gpurun_out/r2t_c3f_single.log
