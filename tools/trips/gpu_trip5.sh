#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "not FiBiNet and not fibinet" > gpurun_out/t5_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t5_tests.log
timeout 300 python tools/exp_gather.py > gpurun_out/t5_exp_gather.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/t5_bench.log 2> gpurun_out/t5_bench.err
echo "bench exit $?"
tail -12 gpurun_out/t5_tests.log | cut -c1-250; cat gpurun_out/t5_exp_gather.log; cat gpurun_out/t5_bench.log | cut -c1-1500; tail -3 gpurun_out/t5_bench.err | cut -c1-300
