#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/mp_profile_sharded.py > gpurun_out/t10_prof2.log 2>&1
echo "prof2 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/t10_bench2.log 2> gpurun_out/t10_bench2.err
echo "bench2 exit $?"
timeout 200 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t10_mg.log 2>&1
echo "mg exit $?"
grep SHARDED_PROFILE gpurun_out/t10_prof2.log; tail -3 gpurun_out/t10_prof2.log | cut -c1-300; cat gpurun_out/t10_bench2.log | cut -c1-1500; grep -v Warn gpurun_out/t10_bench2.err | tail -5 | cut -c1-300; tail -3 gpurun_out/t10_mg.log
