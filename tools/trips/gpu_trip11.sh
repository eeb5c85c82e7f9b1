#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t11_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t11_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t11_bench1.log 2> gpurun_out/t11_bench1.err
echo "bench1 exit $?"
bash tools/gpu_profile_all.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t11_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t11_ncu_bench.log 2>&1
tail -8 gpurun_out/t11_tests.log | cut -c1-300; cat gpurun_out/t11_bench1.log | cut -c1-2600
