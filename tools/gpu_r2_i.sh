#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2i_pytest.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_level_gemm -s 60 -c 1 -f -o gpurun_out/r2i_c5 python tools/prof_r2.py c5 > gpurun_out/r2i_c5.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -30 gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_bench.err
