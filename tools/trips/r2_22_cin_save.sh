#!/bin/bash
# round 2: CIN forward keeps X_k for backward (no recompute), pooled sums by a transposing reduction: parity + bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cin or xDeepFM or xdeepfm or ring" ) > gpurun_out/r2_22_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_22_tests.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_22_tests.log | tail -12 | cut -c1-300
for sv in 1 0; do
RPB_CIN_SAVE_X=$sv timeout 600 python bench.py --workload xdeepfm --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_22_bench_xdeepfm_save$sv.json 2> gpurun_out/r2_22_bench_xdeepfm_save$sv.err
python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_22_bench_xdeepfm_save$sv.json') if l.startswith('{')][-1]
    print('save_x=$sv xdeepfm ms/step', round(j['ms_per_step'],4), 'e2e ms', round(j['e2e']['ms_per_step'],4), 'train_step', j.get('train_step',{}).get('ms_per_step'), 'train_model', j.get('train_model',{}).get('ms_per_step'))
except Exception as e:
    print('no line', e)
PY
tail -2 gpurun_out/r2_22_bench_xdeepfm_save$sv.err | cut -c1-300
done
