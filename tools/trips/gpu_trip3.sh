#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "not FiBiNet and not fibinet" > gpurun_out/t3_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t3_tests.log
timeout 600 python tools/exp_gather.py > gpurun_out/t3_exp_gather.log 2>&1
echo "exp exit $?"
timeout 600 python tools/microbench.py > gpurun_out/t3_micro.log 2>&1
echo "microbench exit $?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/t3_bench.log 2> gpurun_out/t3_bench.err
echo "bench exit $?"
tail -12 gpurun_out/t3_tests.log; cat gpurun_out/t3_exp_gather.log; grep -A3 "linear\|deepfm" gpurun_out/t3_micro.log | head -60; cat gpurun_out/t3_bench.log
