#!/bin/bash
# round 2 (2 GPUs): new pipeline tests, multi-GPU parity (fused core on shards, LR shards, optimizer + re-zero, local init),
# bench at N = 2 with and without the fused core, full default bench at N = 1
mkdir -p gpurun_out profiles
N=${N:-2}
( timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_06_pipeline.log 2>&1
echo "pipeline exit $?" >> gpurun_out/r2_06_pipeline.log; grep -E "passed|failed|Error" gpurun_out/r2_06_pipeline.log | tail -8 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tests/mp_sharded_check.py > gpurun_out/r2_06_sharded_check_n$N.log 2>&1
echo "check rc $?" >> gpurun_out/r2_06_sharded_check_n$N.log
grep -v Warning gpurun_out/r2_06_sharded_check_n$N.log | tail -12 | cut -c1-300
for fused in 1 0; do
  RPB_SHARDED_FUSED=$fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_06_bench${N}_fused$fused.json 2> gpurun_out/r2_06_bench${N}_fused$fused.err
  echo "fused=$fused rc $?"; python -c "import sys,json; j=json.loads(open('gpurun_out/r2_06_bench${N}_fused$fused.json').readline()); print('N=$N fused=$fused ms/step', round(j['ms_per_step'],4), 'value', round(j['value']/1e6,1), 'M/s windows', j['run']['window_ms'], 'e2e', round(j['e2e']['value']/1e6,1))"; grep -v Warning gpurun_out/r2_06_bench${N}_fused$fused.err | tail -3 | cut -c1-300
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_06_bench1.json 2> gpurun_out/r2_06_bench1.err
echo "bench1 rc $?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_06_bench1.json').readline())
for k in ('value','ms_per_step','gpu_launches'): print(k, j[k])
for k in ('e2e','train_step','train_model','zipf_ids','torch_eager_gpu_baseline','cpu_baseline'): print(k, {a:b for a,b in (j.get(k) or {}).items() if a in ('value','ms_per_step','cores','kind','error','sample')})
print('roofline', {a:b for a,b in j['roofline'].items() if a in ('frac','us_per_launch','error')})
PY
grep -v Warning gpurun_out/r2_06_bench1.err | tail -3 | cut -c1-300
