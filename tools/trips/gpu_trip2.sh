#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "not FiBiNet and not fibinet" > gpurun_out/t2_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t2_tests.log
timeout 600 python tools/microbench.py > gpurun_out/t2_micro.log 2>&1
echo "microbench exit $?" | tee -a gpurun_out/t2_micro.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/t2_bench.log 2> gpurun_out/t2_bench.err
echo "bench exit $?" | tee -a gpurun_out/t2_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/t2_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t2_ncu_bench.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_fwd_kernel -s 2 -c 2 -o gpurun_out/t2_gather python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t2_ncu_gather.log 2>&1
echo "ncu gather exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 1 -c 2 -o gpurun_out/t2_gemm python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t2_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
tail -15 gpurun_out/t2_tests.log; tail -60 gpurun_out/t2_micro.log; cat gpurun_out/t2_bench.log; tail -5 gpurun_out/t2_bench.err
