#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_tower_gpu.py tests/test_models_gpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t32_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t32_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/t32_bench.log 2> gpurun_out/t32_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t32_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step > gpurun_out/t32_ncu_bench.log 2>&1
tail -4 gpurun_out/t32_tests.log | cut -c1-300; python -c "
import json; j=json.loads(open('gpurun_out/t32_bench.log').readline()); print('ms', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['value'])"
