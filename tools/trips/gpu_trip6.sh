#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/t6_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t6_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t6_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/t6_bench2.log 2> gpurun_out/t6_bench2.err
echo "bench2 exit $?"
tail -14 gpurun_out/t6_tests.log | cut -c1-300; cat gpurun_out/t6_bench2.log | cut -c1-1200; tail -12 gpurun_out/t6_bench2.err | cut -c1-400
