#!/bin/bash
mkdir -p gpurun_out
tools/bin/ubench_store > gpurun_out/ubench_store.log 2>&1
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_c8.log
timeout 600 python tools/qt_sweep.py "" "B200_CHAIN_K=2" "B200_CHAIN_K=8" "B200_CHAIN_CTAS=4" "B200_CHAIN_CTAS=2" "B200_CHAIN_CTAS=4,B200_CHAIN_K=2" "B200_DBG=2" "B200_DBG=1" "B200_CHAIN_PROF=1" > gpurun_out/qt_c8.log 2>&1
cat gpurun_out/ubench_store.log; tail -n 6 gpurun_out/pytest_c8.log; grep -v "^\[chain prof\]" gpurun_out/qt_c8.log; grep "chain prof" gpurun_out/qt_c8.log | tail -2; grep "accum prof" gpurun_out/qt_c8.log | tail -1
