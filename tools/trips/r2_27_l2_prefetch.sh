#!/bin/bash
# round 2: L2 prefetch warp in the one-kernel DeepFM forward: parity of the fused tests, bench with 0 / 1 / 2 / 4 tiles ahead
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "fused or one_kernel or DeepFM or deepfm" ) > gpurun_out/r2_27_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_27_tests.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_27_tests.log | tail -6 | cut -c1-300
for pf in 0 1 2 4; do
  RPB_OPTIONS=fused_l2_prefetch=$pf timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_27_bench_pf$pf.json 2> gpurun_out/r2_27_bench_pf$pf.err
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_27_bench_pf$pf.json') if l.startswith('{')][-1]
    print('l2_prefetch=$pf ms/step', round(j['ms_per_step'],4), 'value M/s', round(j['value']/1e6,2), 'fwd us', round(j['roofline']['us_per_launch'],2), 'frac', round(j['roofline']['frac'],4), 'e2e loss', j['e2e']['loss'])
except Exception as e:
    print('no line', e)
PY
done
