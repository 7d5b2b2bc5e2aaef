#!/bin/bash
# factored kernels: parity tests + config 3 timing (dense level path vs factor programs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -q -x -k "factored or config3 or config5" 2>&1 | tail -2
REPS=3 timeout 200 python tools/prof_r2.py c3f 50000 2>&1 | tail -2
