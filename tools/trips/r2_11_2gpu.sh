#!/bin/bash
# round 2 (2 GPUs): flat gradient shards + multi-tensor bucket, parity at N = 2, fetch-warp count at N = 2, exact-lazy Adam test,
# full 1-GPU suite
mkdir -p gpurun_out
N=${N:-2}
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_11_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_11_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_11_tests.log | tail -12 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tests/mp_sharded_check.py > gpurun_out/r2_11_sharded_check_n$N.log 2>&1
echo "check rc $?" >> gpurun_out/r2_11_sharded_check_n$N.log
grep -v "Warning\|\*\*\*\|OMP_NUM\|Successfully set" gpurun_out/r2_11_sharded_check_n$N.log | tail -7 | cut -c1-300
for nf in 8 16; do
  RPB_OPTIONS=fused_fetch_warps=$nf timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_11_bench${N}_nf$nf.json 2> gpurun_out/r2_11_bench${N}_nf$nf.err
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_11_bench${N}_nf$nf.json') if l.startswith('{')][-1]
    print('N=$N fetch_warps=$nf ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,1), 'windows', j['run']['window_ms'], 'launches/step', j['gpu_launches']//j['steps'])
except Exception as e:
    print('no line', e)
PY
  grep -v "Warning\|run_backward\|\*\*\*\|OMP_NUM" gpurun_out/r2_11_bench${N}_nf$nf.err | tail -3 | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 \
    tools/mp_profile_sharded.py > gpurun_out/r2_11_profile_n$N.log 2>&1
grep -A16 SHARDED_PROFILE gpurun_out/r2_11_profile_n$N.log | cut -c1-200
