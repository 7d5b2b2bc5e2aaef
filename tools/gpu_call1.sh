#!/bin/bash
# dev script: one gpurun call = synthetic tests of new kernels, timings, full gpu suite, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_synthetic.py -q 2>&1 | tail -40 > gpurun_out/pytest_synth.log
timeout 300 python tools/quick_time.py c2_full_layout 5 > gpurun_out/qt.log 2>&1
B200_CHAIN_CTAS=6 timeout 300 python tools/quick_time.py c2_full_layout 5 > gpurun_out/qt_ctas6.log 2>&1
timeout 600 python tools/time_big_d.py > gpurun_out/bigd.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_synthetic.py 2>&1 | tail -25 > gpurun_out/pytest_rest.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/pytest_synth.log gpurun_out/qt.log gpurun_out/bigd.log gpurun_out/pytest_rest.log gpurun_out/bench.log
