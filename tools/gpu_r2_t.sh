#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "factored" 2>&1 | tail -25 > gpurun_out/r2t_pytest.log
tail -5 gpurun_out/r2t_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3f 50000 > gpurun_out/r2t_c3f.log 2>&1; tail -3 gpurun_out/r2t_c3f.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fj64_ -c 2 -o gpurun_out/r2t_fj64 python tools/prof_r2.py c3f 10000 > gpurun_out/r2t_ncu.log 2>&1
tail -2 gpurun_out/r2t_ncu.log
