#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2l_pytest.log
timeout 200 python tools/prof_fac.py > gpurun_out/r2l_fac.log 2>&1
tail -4 gpurun_out/r2l_pytest.log; grep -v "^==" gpurun_out/r2l_fac.log | tail -3
