#!/bin/bash
# round 2: CIN kernels after the pipelined wgrad loader: parity, bench, ncu --set full of one step's CIN launches
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cin or xDeepFM or xdeepfm" ) > gpurun_out/r2_18_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_18_tests.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_18_tests.log | tail -12 | cut -c1-300
timeout 600 python bench.py --workload xdeepfm --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_18_bench.json 2> gpurun_out/r2_18_bench.err
python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_18_bench.json') if l.startswith('{')][-1]
    print('ms/step', round(j['ms_per_step'],3), 'e2e ms', j['e2e']['ms_per_step'], 'train_step', j.get('train_step',{}).get('ms_per_step'), 'train_model', j.get('train_model',{}).get('ms_per_step'))
except Exception as e:
    print('no line', e)
PY
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'cin_' --launch-skip 12 --launch-count 12 -f -o gpurun_out/r2_18_cin \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_18_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/r2_18_cin.ncu-rep 2>&1 | tail -14
