#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_c11.log
timeout 600 python tools/qt_sweep.py "" "B200_CHAIN_EARLY=0" "B200_CHAIN_K=2" "B200_CHAIN_SLEEP=100" "B200_CHAIN_SLEEP=250" "B200_CHAIN_CTAS=3" "B200_CHAIN_CTAS=2" "B200_TRIE_LEX=1" "B200_TRIE_LEX=1,B200_CHAIN_K=2" "B200_ACC_RSUB=8" "B200_CHAIN_PROF=1" > gpurun_out/qt_c11.log 2>&1
tail -n 3 gpurun_out/pytest_c11.log; grep -v "prof\]" gpurun_out/qt_c11.log; grep "chain prof" gpurun_out/qt_c11.log | tail -2; grep "accum prof" gpurun_out/qt_c11.log | tail -1
