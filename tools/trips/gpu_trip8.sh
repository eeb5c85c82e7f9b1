#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t8_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t8_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t8_bench1.log 2> gpurun_out/t8_bench1.err
echo "bench1 exit $?"
timeout 300 python tools/microbench.py > gpurun_out/t8_micro.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t8_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t8_ncu_bench.log 2>&1
echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 2 -c 4 -o gpurun_out/t8_gemm python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t8_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tf32x3_kernel -s 0 -c 2 -o gpurun_out/t8_wgrad python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t8_ncu_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_bwd_kernel -s 0 -c 1 -o gpurun_out/t8_scatter python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t8_ncu_scatter.log 2>&1
tail -14 gpurun_out/t8_tests.log | cut -c1-300; cat gpurun_out/t8_bench1.log | cut -c1-700
