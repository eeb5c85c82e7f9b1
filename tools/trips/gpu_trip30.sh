#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t30_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t30_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/t30_bench.log 2> gpurun_out/t30_bench.err
tail -6 gpurun_out/t30_tests.log | cut -c1-400; cut -c1-300 gpurun_out/t30_bench.log
