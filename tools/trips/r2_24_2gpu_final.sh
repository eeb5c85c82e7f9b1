#!/bin/bash
# round 2, final 2-GPU pass: parity script, DeepFM and xDeepFM bench lines at N = 2
mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    tests/mp_sharded_check.py > gpurun_out/r2_24_sharded_check_n$N.log 2>&1
echo "check rc $?" >> gpurun_out/r2_24_sharded_check_n$N.log
grep -v "Warning\|\*\*\*\|OMP_NUM\|Successfully set" gpurun_out/r2_24_sharded_check_n$N.log | tail -8 | cut -c1-300
for wl in deepfm xdeepfm; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 \
      bench.py --gpus $N --steps 50 --warmup 5 --workload $wl > gpurun_out/r2_24_bench${N}_$wl.json 2> gpurun_out/r2_24_bench${N}_$wl.err
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_24_bench${N}_$wl.json') if l.startswith('{')][-1]
    print('$wl N=$N ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,1), 'e2e M/s', round(j['e2e']['value']/1e6,1), 'windows', j['run']['window_ms'])
except Exception as e:
    print('no line', e)
PY
  grep -v "Warning\|run_backward\|\*\*\*\|OMP_NUM\|AccumulateGrad" gpurun_out/r2_24_bench${N}_$wl.err | tail -3 | cut -c1-300
done
