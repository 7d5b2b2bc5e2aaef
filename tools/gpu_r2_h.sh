#!/bin/bash
# two B200s: hardware tests of the multi-GPU paths (NCCL all-gather, fused peer-store exchange, single-process concurrent) + bench N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v -s 2>&1 | tail -25 > gpurun_out/r2h_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -12 gpurun_out/r2h_pytest_multi.log; cat gpurun_out/r2h_bench_n2.json | cut -c1-6000; tail -5 gpurun_out/r2h_bench_n2.err | cut -c1-300
