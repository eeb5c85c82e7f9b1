#!/bin/bash
# round 2 (2 GPUs): config 5 (MMOE, 8 x 100M-row tables row-sharded) at N = 2, xDeepFM at N = 2 (sharded LR tables), and on
# one GPU: the tests touched since the last full run, the default bench (train_model leg), xDeepFM train_step
mkdir -p gpurun_out
N=${N:-2}
( timeout 900 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider -k "aitm or masknet or lr or adam or train_model or stager or dropout" ) > gpurun_out/r2_08_tests.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_08_tests.log | tail -8 | cut -c1-300
for wl in mmoe_cfg5 xdeepfm; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
      bench.py --gpus $N --workload $wl --steps 20 --warmup 5 > gpurun_out/r2_08_bench${N}_$wl.json 2> gpurun_out/r2_08_bench${N}_$wl.err
  echo "$wl N=$N rc $?"
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_08_bench${N}_$wl.json') if l.startswith('{')][-1]
    print('$wl N=$N ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,2), 'windows', j['run']['window_ms'], 'e2e M/s', round(j['e2e']['value']/1e6,2), 'tables/GPU GB', round(j['run']['tables_per_gpu_bytes']/1e9,1))
except Exception as e:
    print('$wl: no line', e)
PY
  grep -v "Warning\|run_backward\|\*\*\*\|OMP_NUM" gpurun_out/r2_08_bench${N}_$wl.err | tail -4 | cut -c1-300
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_08_bench1.json 2> gpurun_out/r2_08_bench1.err
timeout 600 python bench.py --workload xdeepfm --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_08_bench1_xdeepfm.json 2> gpurun_out/r2_08_bench1_xdeepfm.err
python - <<'PY'
import json
for f in ('r2_08_bench1','r2_08_bench1_xdeepfm'):
    try:
        j=[json.loads(l) for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1]
        print(f, 'ms/step', round(j['ms_per_step'],4)); 
        for k in ('train_step','train_model'): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in (j.get(k) or {}).items() if a in ('value','ms_per_step','error')})
    except Exception as e:
        print(f, 'no line', e)
PY
