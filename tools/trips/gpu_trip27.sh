#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t27_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t27_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/t27_bench.log 2> gpurun_out/t27_bench.err
RPB_PARALLEL_WGRAD=0 timeout 300 python bench.py --no-cpu-baseline --no-train-step > gpurun_out/t27_bench_serial.log 2>&1
tail -6 gpurun_out/t27_tests.log | cut -c1-400; cat gpurun_out/t27_bench.log; tail -3 gpurun_out/t27_bench.err | cut -c1-300; cut -c1-300 gpurun_out/t27_bench_serial.log
