#!/bin/bash
# round 2: default bench line after the train_model-leg change (median of 3 calls)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_29_bench_deepfm.json 2> gpurun_out/r2_29_bench_deepfm.err
python - <<PY
import json
j=[json.loads(l) for l in open('gpurun_out/r2_29_bench_deepfm.json') if l.startswith('{')][-1]
print('ms/step', round(j['ms_per_step'],4), 'value', round(j['value']/1e6,2), 'e2e', round(j['e2e']['value']/1e6,2), 'roofline', round(j['roofline']['frac'],4), 'train_step', j['train_step']['ms_per_step'], 'train_model', j['train_model']['ms_per_step'], j['train_model'].get('runs_ms_per_step'), 'cpu', j['cpu_baseline']['value'], 'clocks', j['clocks'])
PY
