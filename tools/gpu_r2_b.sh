#!/bin/bash
# round 2, call B (two B200s): the 2-GPU NCCL parity test on hardware + the strong-scaling bench at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2b_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2b_gpu.txt 2>&1
df -h /dev/shm >> gpurun_out/r2b_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v 2>&1 | tail -15 > gpurun_out/r2b_pytest_multi.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
grep -i "NVLS\|via P2P\|Channel 00" gpurun_out/r2b_bench_n2.err | head -8 > gpurun_out/r2b_nccl_info.txt
cat gpurun_out/r2b_pytest_multi.log | tail -8; cat gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err | cut -c1-300
