#!/bin/bash
# round 2, last call: float4 dropout kernels — full GPU suite and the xDeepFM / AutoInt bench lines
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_31_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_31_tests.log
grep -E "passed|failed|FAILED|Error|assert|exit" gpurun_out/r2_31_tests.log | tail -8 | cut -c1-300
for wl in xdeepfm autoint; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-train-step > gpurun_out/r2_31_bench_$wl.json 2> gpurun_out/r2_31_bench_$wl.err
python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_31_bench_$wl.json') if l.startswith('{')][-1]
    print('$wl ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,2), 'e2e M/s', round(j['e2e']['value']/1e6,2))
except Exception as e:
    print('no line', e)
PY
done
