#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t24_tower.log 2>&1
echo "tower exit $?" >> gpurun_out/t24_tower.log
( timeout 600 python -m pytest tests/test_models_gpu.py tests/test_pipeline_gpu.py tests/test_zz_linear_tc_gpu.py -m gpu -q --maxfail=5 -p no:cacheprovider ) > gpurun_out/t24_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t24_tests.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/t24_bench.log 2> gpurun_out/t24_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t24_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t24_ncu_bench.log 2>&1
tail -12 gpurun_out/t24_tower.log | cut -c1-300; tail -5 gpurun_out/t24_tests.log | cut -c1-300; cut -c1-400 gpurun_out/t24_bench.log; tail -3 gpurun_out/t24_bench.err | cut -c1-300
