#!/bin/bash
# round 2: warp-uniform MMA issue in wgrad_tf32x3 / gemm_tf32x3_v2 and the CIN kernels, two-warp wgrad loader: full suite, both benches
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_19_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_19_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_19_tests.log | tail -8 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_19_bench.json 2> gpurun_out/r2_19_bench.err
timeout 600 python bench.py --workload xdeepfm --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_19_bench_xdeepfm.json 2> gpurun_out/r2_19_bench_xdeepfm.err
python - <<PY
import json
for f in ('gpurun_out/r2_19_bench.json', 'gpurun_out/r2_19_bench_xdeepfm.json'):
    try:
        j=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
        print(j['config']['workload'] if 'workload' in j['config'] else '', 'ms/step', round(j['ms_per_step'],4), 'value', round(j['value']/1e6,2), 'e2e ms', round(j['e2e']['ms_per_step'],4), 'roofline', round(j['roofline']['frac'],4), j['roofline']['us_per_launch'], 'train_step', j.get('train_step',{}).get('ms_per_step'), 'train_model', j.get('train_model',{}).get('ms_per_step'))
    except Exception as e:
        print('no line', f, e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_19_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_19_ncu_bench.log 2>&1
python tools/step_list.py gpurun_out/r2_19_bench_launches.csv | cut -c1-60,100-170
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_19_xdeepfm_launches.csv \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_19_ncu_bench2.log 2>&1
python tools/step_list.py gpurun_out/r2_19_xdeepfm_launches.csv 2>&1 | grep -i "cin\|total" | cut -c1-60,100-170
