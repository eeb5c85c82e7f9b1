#!/bin/bash
# final state of round 1: full GPU suite, smoke, bench (both arms), launch list, ncu full of the fused forward kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t45_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t45_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/t45_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/t45_bench.json 2> gpurun_out/t45_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t45_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/t45_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:deepfm_fwd_fused --launch-skip 3 --launch-count 1 -o gpurun_out/t45_fused -f python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/t45_ncu_full.log 2>&1
tail -4 gpurun_out/t45_tests.log | cut -c1-300; tail -1 gpurun_out/t45_smoke.log | cut -c1-300
python - <<'PY'
import json
j=json.loads(open('gpurun_out/t45_bench.json').readline())
for k in ('value','ms_per_step','gpu_launches'): print(k, j[k])
for k in ('e2e','train_step','zipf_ids','torch_eager_gpu_baseline','cpu_baseline'): print(k, {a:b for a,b in j[k].items() if a in ('value','ms_per_step','cores')})
print('roofline frac', j['roofline']['frac'], j['roofline']['no_materialise']['frac'])
PY
