#!/bin/bash
# eight B200s: the strong-scaling bench at N = 8 (what the driver runs at round end), NVLS / topology info
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2n8_topo.txt 2>&1
NCCL_DEBUG=INFO timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err
grep -i "NVLS\|nvls" gpurun_out/r2n8_bench.err | head -5 > gpurun_out/r2n8_nccl.txt
cat gpurun_out/r2n8_bench.json | cut -c1-3000; grep -v "NCCL INFO" gpurun_out/r2n8_bench.err | tail -8 | cut -c1-300
