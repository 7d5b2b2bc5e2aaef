#!/bin/bash
# quick head check: smoke + the drop-in tests that run the factored / Lindblad device paths through pyGSTi
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-260
timeout 90 python -m pytest tests/test_gpu_pygsti_dropin.py -m gpu -q -x -k "three_qubit or four_qubit or lindblad" 2>&1 | tail -2
