#!/bin/bash
# round 2: wgrad operand warps software-pipelined, backward G warps prefetch the next tile's rows into L2: parity, bench, launch list
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cin or xDeepFM or xdeepfm" ) > gpurun_out/r2_25_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_25_tests.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_25_tests.log | tail -12 | cut -c1-300
timeout 600 python bench.py --workload xdeepfm --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_25_bench.json 2> gpurun_out/r2_25_bench.err
python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_25_bench.json') if l.startswith('{')][-1]
    print('ms/step', round(j['ms_per_step'],4), 'e2e ms', j['e2e']['ms_per_step'], 'train_step', j.get('train_step',{}).get('ms_per_step'), 'train_model', j.get('train_model',{}).get('ms_per_step'))
except Exception as e:
    print('no line', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_25_xdeepfm_launches.csv \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_25_ncu.log 2>&1
python tools/step_list.py gpurun_out/r2_25_xdeepfm_launches.csv 2>&1 | grep -i "cin_.*tc_kernel\|total" | cut -c1-60,100-170
