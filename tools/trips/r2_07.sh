#!/bin/bash
# round 2 (1 GPU): full GPU suite, smoke, default bench (all legs), bench --workload xdeepfm / autoint
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_07_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_07_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_07_tests.log | tail -15 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_07_bench_deepfm.json 2> gpurun_out/r2_07_bench_deepfm.err
for wl in xdeepfm autoint; do
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/r2_07_bench_$wl.json 2> gpurun_out/r2_07_bench_$wl.err
  echo "$wl rc $?"
done
python - <<'PY'
import json
for wl in ('deepfm','xdeepfm','autoint'):
    try:
        j=[json.loads(l) for l in open(f'gpurun_out/r2_07_bench_{wl}.json') if l.startswith('{')][-1]
    except Exception as e:
        print(wl, 'no line', e); continue
    print(wl, 'ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,2))
    for k in ('e2e','train_step','train_model','zipf_ids','torch_eager_gpu_baseline','cpu_baseline'): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in (j.get(k) or {}).items() if a in ('value','ms_per_step','cores','kind','error','skipped')})
    print('    roofline', {a:b for a,b in (j.get('roofline') or {}).items() if a in ('frac','us_per_launch','error','kernel')})
PY
for wl in deepfm xdeepfm autoint; do grep -v "Warning\|run_backward\|INFO" gpurun_out/r2_07_bench_$wl.err | tail -3 | cut -c1-300; done
