#!/bin/bash
# The round-end check that was run under gpurun (one B200): full GPU test suite, both bench arms, the ncu launch list of the
# bench command and one ncu --set full capture of the dominant kernel.  Outputs go to gpurun_out/ (scratch); the summaries
# that are meant to be judged were copied to profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_final.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_final.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accum_trie -s 3 -c 1 -f -o gpurun_out/r01_accum_final2 python tools/qt_sweep.py "" > gpurun_out/ncu_final2.log 2>&1
tail -n 3 gpurun_out/pytest_final.log; cat gpurun_out/bench_ref_final.json | cut -c1-220; cat gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
