#!/bin/bash
# full GPU suite + smoke + bench (both arms): state check after re-entry
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -p no:cacheprovider ) > gpurun_out/t21_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/t21_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t21_smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/t21_bench.log 2> gpurun_out/t21_bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/t21_ref.log 2> gpurun_out/t21_ref.err
nproc > gpurun_out/t21_host.txt; free -g >> gpurun_out/t21_host.txt; nvidia-smi -L >> gpurun_out/t21_host.txt
tail -25 gpurun_out/t21_tests.log | cut -c1-250; tail -3 gpurun_out/t21_smoke.log | cut -c1-300; cut -c1-1500 gpurun_out/t21_bench.log; tail -5 gpurun_out/t21_bench.err | cut -c1-300; cut -c1-600 gpurun_out/t21_ref.log
