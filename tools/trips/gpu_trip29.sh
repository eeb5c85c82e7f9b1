#!/bin/bash
# 2 GPUs: sharded parity + bench at N=2 (driver-style torchrun launch)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t29_mp.log 2>&1
echo "mp exit $?" >> gpurun_out/t29_mp.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 ) > gpurun_out/t29_bench2.log 2> gpurun_out/t29_bench2.err
tail -15 gpurun_out/t29_mp.log | cut -c1-600; cat gpurun_out/t29_bench2.log | cut -c1-1500; tail -5 gpurun_out/t29_bench2.err | cut -c1-300
