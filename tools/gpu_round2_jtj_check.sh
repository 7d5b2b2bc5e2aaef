#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "jtj" 2>&1 | tail -2
timeout 120 python tools/oz_test.py c2_full_layout 4 2>&1 | head -3
