#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tower_gpu.py tests/test_models_gpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t23_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t23_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/t23_bench.log 2> gpurun_out/t23_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t23_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t23_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tower_tail --launch-skip 6 --launch-count 2 -o gpurun_out/t23_tower -f python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t23_ncu_full.log 2>&1
tail -4 gpurun_out/t23_tests.log | cut -c1-300; cut -c1-400 gpurun_out/t23_bench.log; tail -3 gpurun_out/t23_bench.err | cut -c1-300; tail -3 gpurun_out/t23_ncu_full.log
