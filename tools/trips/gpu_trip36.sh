#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t36_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t36_tests.log
( time timeout 600 python bench.py ) > gpurun_out/t36_bench.json 2> gpurun_out/t36_bench.err
tail -5 gpurun_out/t36_tests.log | cut -c1-300; tail -4 gpurun_out/t36_bench.err | cut -c1-200
python - <<'PY'
import json
j=json.loads(open('gpurun_out/t36_bench.json').readline())
for k in ('value','ms_per_step','gpu_launches','clocks'): print(k, j[k])
for k in ('e2e','train_step','zipf_ids','torch_eager_gpu_baseline','cpu_baseline'): print(k, j.get(k))
print('roofline frac', j['roofline']['frac'], j['roofline']['no_materialise']['frac'])
PY
