#!/bin/bash
# tower tail kernels: parity + bench + launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t22_tower.log 2>&1
echo "tower exit $?" >> gpurun_out/t22_tower.log
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider ) > gpurun_out/t22_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t22_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/t22_bench.log 2> gpurun_out/t22_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t22_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t22_ncu_bench.log 2>&1
tail -15 gpurun_out/t22_tower.log | cut -c1-300; tail -8 gpurun_out/t22_tests.log | cut -c1-300; cut -c1-1200 gpurun_out/t22_bench.log; tail -5 gpurun_out/t22_bench.err | cut -c1-300
