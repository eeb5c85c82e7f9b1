#!/bin/bash
# round 2: zero_grad opening the step, re-zero co-resident with the forward kernel (GraphedStep.ring): parity test, bench A/B,
# CIN pooled transposing reduction (xDeepFM bench), launch check
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "ring or pipeline" ) > gpurun_out/r2_20_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_20_tests.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_20_tests.log | tail -12 | cut -c1-300
timeout 300 python tools/exp/overlap_zero.py 148 296 2>&1 | grep -v Warning | tail -5
for zf in 1 0; do
  timeout 600 python bench.py --zero-first $zf --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_20_bench_zf$zf.json 2> gpurun_out/r2_20_bench_zf$zf.err
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_20_bench_zf$zf.json') if l.startswith('{')][-1]
    print('zero_first=$zf ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,2), 'e2e ms', round(j['e2e']['ms_per_step'],4), 'e2e loss', j['e2e']['loss'], 'launches/step', j['gpu_launches']//j['steps'])
except Exception as e:
    print('no line', e)
PY
  tail -2 gpurun_out/r2_20_bench_zf$zf.err | cut -c1-300
done
