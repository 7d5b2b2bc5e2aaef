#!/bin/bash
# ncu --set full of the small late-round kernels (factor-program probabilities, Lindblad recursion) out of one bench pass
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:k_probs_fdmma|k_lind_dexp16|k_lind_gen16" -c 4 -o gpurun_out/r2_small python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_small.log 2>&1
ls -la gpurun_out/r2_small.ncu-rep
