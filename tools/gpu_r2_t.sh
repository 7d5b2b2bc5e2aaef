#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "factored or config3" 2>&1 | tail -25 > gpurun_out/r2t_pytest.log
tail -5 gpurun_out/r2t_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3f 50000 > gpurun_out/r2t_c3f.log 2>&1; tail -3 gpurun_out/r2t_c3f.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches_c3f.csv python tools/prof_r2.py c3f 50000 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2t_launches_c3f.csv 2>/dev/null | grep "k_fj"
