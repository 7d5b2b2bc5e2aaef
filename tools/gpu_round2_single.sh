#!/bin/bash
# round 2, one B200 (as run under gpurun): : smoke, full -m gpu suite, bench (both arms), launch list of the bench command, ncu captures of the new kernels
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2z_pytest.log; tail -2 gpurun_out/r2z_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2z_clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -2 gpurun_out/r2z_bench.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; tail -c 600 gpurun_out/r2z_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2z_launches_bench.csv 2>/dev/null | head -14
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_oz_syrk -c 1 -o gpurun_out/r2z_oz_syrk python tools/oz_one.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fj64_backward -c 1 -o gpurun_out/r2z_fj64_backward python tools/prof_r2.py c3f 10000 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_accum_trie_d16 -c 1 -o gpurun_out/r2z_accum python tools/quick_time.py c2_full_layout 1 > /dev/null 2>&1
ls -la gpurun_out/r2z_*
