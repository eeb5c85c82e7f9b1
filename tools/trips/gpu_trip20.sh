#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_linear_tc_gpu.py tests/test_models_gpu.py -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t20_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t20_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t20_bench.log 2> gpurun_out/t20_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t20_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t20_ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rpb --csv --log-file gpurun_out/t20_mmoe_launches.csv python tools/profile_all.py --model MMOE --steps 2 > gpurun_out/t20_mmoe.log 2>&1
tail -8 gpurun_out/t20_tests.log | cut -c1-300; cut -c1-300 gpurun_out/t20_bench.log; tail -3 gpurun_out/t20_bench.err | cut -c1-300
