#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2k_pytest.log
REPS=3 timeout 300 python tools/prof_r2.py c3 50000 > gpurun_out/r2k_c3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_probs_factored -s 1 -c 1 -f -o gpurun_out/r2k_fac python tools/prof_fac.py > gpurun_out/r2k_fac.log 2>&1
timeout 200 python tools/prof_fac.py > gpurun_out/r2k_fac2.log 2>&1
tail -4 gpurun_out/r2k_pytest.log; grep -v "^==" gpurun_out/r2k_c3.log gpurun_out/r2k_fac2.log | tail -5
