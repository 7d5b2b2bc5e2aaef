#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "factored" 2>&1 | tail -25 > gpurun_out/r2r_pytest.log
tail -25 gpurun_out/r2r_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3f 50000 > gpurun_out/r2r_c3f.log 2>&1; tail -5 gpurun_out/r2r_c3f.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_launches_c3f.csv python tools/prof_r2.py c3f 10000 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2r_launches_c3f.csv 2>/dev/null | grep -v "k_level_gemm\|k_level_init" | tail -12
