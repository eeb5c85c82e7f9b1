#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t7_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t7_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/t7_bench2.log 2> gpurun_out/t7_bench2.err
echo "bench2 exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t7_bench1.log 2> gpurun_out/t7_bench1.err
echo "bench1 exit $?"
timeout 300 python tools/microbench.py > gpurun_out/t7_micro.log 2>&1
tail -14 gpurun_out/t7_tests.log | cut -c1-300; cat gpurun_out/t7_bench2.log | cut -c1-1300; grep -v Warning gpurun_out/t7_bench2.err | tail -12 | cut -c1-300; cat gpurun_out/t7_bench1.log | cut -c1-900; grep -A3 "linear1_dw\|scatter\|deepfm_fwdbwd" gpurun_out/t7_micro.log | head -30
