#!/bin/bash
# NOT RUN YET (written at the end of round 1 with the GPU budget spent): first thing to run in round 2 on 2 GPUs
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/trips/r2_sharded_fused.sh'
# 1. parity of the fused DeepFM core on row-sharded tables (ops.SHARDED_FUSED) against the single-GPU step on the global batch,
# 2. bench at N = 2 with the separate sharded kernels (default) and with the fused core.
mkdir -p gpurun_out
N=${N:-2}
RPB_SHARDED_FUSED=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tests/mp_sharded_check.py > gpurun_out/r2_sharded_check.log 2>&1
echo "check rc $?" >> gpurun_out/r2_sharded_check.log
grep -v Warning gpurun_out/r2_sharded_check.log | tail -15 | cut -c1-300
for fused in 0 1; do
  RPB_SHARDED_FUSED=$fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_bench${N}_fused$fused.json 2> gpurun_out/r2_bench${N}_fused$fused.err
  echo "fused=$fused rc $?"; cut -c1-400 gpurun_out/r2_bench${N}_fused$fused.json; grep -v Warning gpurun_out/r2_bench${N}_fused$fused.err | tail -4 | cut -c1-300
done
