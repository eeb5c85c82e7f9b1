#!/bin/bash
# round 2 (8 GPUs): multi-GPU parity at N = 8, DeepFM bench + per-kernel table at N = 8, config 5 (MMOE, 8 x 100M-row tables)
mkdir -p gpurun_out
N=${N:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tests/mp_sharded_check.py > gpurun_out/r2_10_sharded_check_n$N.log 2>&1
echo "check rc $?" >> gpurun_out/r2_10_sharded_check_n$N.log
grep -v "Warning\|\*\*\*\|OMP_NUM\|Successfully set" gpurun_out/r2_10_sharded_check_n$N.log | tail -8 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_10_bench${N}_deepfm.json 2> gpurun_out/r2_10_bench${N}_deepfm.err
echo "deepfm rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 \
    tools/mp_profile_sharded.py > gpurun_out/r2_10_profile_n$N.log 2>&1
grep -A30 SHARDED_PROFILE gpurun_out/r2_10_profile_n$N.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N --workload mmoe_cfg5 --steps 20 --warmup 5 > gpurun_out/r2_10_bench${N}_mmoe_cfg5.json 2> gpurun_out/r2_10_bench${N}_mmoe_cfg5.err
echo "mmoe rc $?"
python - <<PY
import json
for wl in ('deepfm','mmoe_cfg5'):
    try:
        j=[json.loads(l) for l in open('gpurun_out/r2_10_bench${N}_%s.json' % wl) if l.startswith('{')][-1]
        print(wl, 'N=$N ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,2), 'windows', j['run']['window_ms'], 'e2e M/s', round(j['e2e']['value']/1e6,2), 'tables/GPU GB', round(j['run']['tables_per_gpu_bytes']/1e9,1))
    except Exception as e:
        print(wl, 'no line', e)
PY
for wl in deepfm mmoe_cfg5; do grep -v "Warning\|run_backward\|\*\*\*\|OMP_NUM" gpurun_out/r2_10_bench${N}_$wl.err | tail -3 | cut -c1-300; done
