#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c22.json 2> gpurun_out/bench_ref_c22.err
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c22.json 2> gpurun_out/bench_c22.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c22.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_c22.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accum_trie -s 3 -c 1 -f -o gpurun_out/r01_accum_final python tools/qt_sweep.py "" > gpurun_out/ncu_c22a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_trie_chains -s 17 -c 1 -f -o gpurun_out/r01_chains_final python tools/qt_sweep.py "" > gpurun_out/ncu_c22b.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_level_gemm -s 200 -c 1 -f -o gpurun_out/r01_c5_level_gemm python tools/time_c5.py > gpurun_out/ncu_c22c.log 2>&1
timeout 300 python tools/time_c5.py > gpurun_out/c5_c22.log 2>&1
cat gpurun_out/bench_ref_c22.json gpurun_out/bench_c22.json; cat gpurun_out/c5_c22.log; ls -la gpurun_out/*.ncu-rep
