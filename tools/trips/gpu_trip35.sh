#!/bin/bash
# 8 GPUs: driver-style launch of both arms
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 100 --warmup 5 ) > gpurun_out/t35_bench8.log 2> gpurun_out/t35_bench8.err
echo "rc $?" >> gpurun_out/t35_bench8.log
cut -c1-1400 gpurun_out/t35_bench8.log; grep -v Warning gpurun_out/t35_bench8.err | tail -6 | cut -c1-300
