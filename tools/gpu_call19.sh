#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_c19.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c19.json 2> gpurun_out/bench_c19.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c19.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_c19.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_accum_trie|k_trie_chains|k_trie_prepare" -s 12 -c 3 -f -o gpurun_out/trie_r01_final python tools/qt_sweep.py "" > gpurun_out/ncu_final_c19.log 2>&1
tail -n 8 gpurun_out/pytest_c19.log; cat gpurun_out/bench_c19.json; tail -3 gpurun_out/bench_c19.err; ls -la gpurun_out/*.ncu-rep
