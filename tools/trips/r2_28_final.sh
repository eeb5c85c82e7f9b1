#!/bin/bash
# round 2, final 1-GPU pass (re-run after the last kernel changes): full suite, smoke, the three bench workloads (full legs on the default one), launch lists, ncu --set full
# of the CIN kernels of one xDeepFM step
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_28_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_28_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_28_tests.log | tail -8 | cut -c1-300
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -2 | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_28_bench_deepfm.json 2> gpurun_out/r2_28_bench_deepfm.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_28_bench_reference.json 2> gpurun_out/r2_28_bench_reference.err
timeout 900 python bench.py --workload xdeepfm --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_28_bench_xdeepfm.json 2> gpurun_out/r2_28_bench_xdeepfm.err
timeout 900 python bench.py --workload autoint --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_28_bench_autoint.json 2> gpurun_out/r2_28_bench_autoint.err
python - <<PY
import json
for f in ('deepfm', 'reference', 'xdeepfm', 'autoint'):
    try:
        j=[json.loads(l) for l in open(f'gpurun_out/r2_28_bench_{f}.json') if l.startswith('{')][-1]
        print(f, 'ms/step', round(j['ms_per_step'],4), 'value', round(j['value']/1e6,3), 'e2e', round(j['e2e']['value']/1e6,3), 'roofline', (j.get('roofline') or {}).get('frac'), 'train_step', (j.get('train_step') or {}).get('ms_per_step'), 'train_model', (j.get('train_model') or {}).get('ms_per_step'), 'eager', (j.get('torch_eager_gpu_baseline') or {}).get('ms_per_step'), 'cpu', (j.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print('no line', f, e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_28_deepfm_launches.csv \
    python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_28_ncu1.log 2>&1
python tools/step_list.py gpurun_out/r2_28_deepfm_launches.csv > gpurun_out/r2_28_deepfm_step.txt; tail -1 gpurun_out/r2_28_deepfm_step.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_28_xdeepfm_launches.csv \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_28_ncu2.log 2>&1
python tools/step_list.py gpurun_out/r2_28_xdeepfm_launches.csv > gpurun_out/r2_28_xdeepfm_step.txt; grep -i "cin_\|total" gpurun_out/r2_28_xdeepfm_step.txt | cut -c1-60,100-170
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'cin_.*tc_kernel' --launch-skip 9 --launch-count 9 -f -o gpurun_out/r2_28_cin \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_28_ncu3.log 2>&1
python tools/ncu_table.py gpurun_out/r2_28_cin.ncu-rep > gpurun_out/r2_28_cin_table.md 2>&1; tail -11 gpurun_out/r2_28_cin_table.md
