#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_linear_tc_gpu.py tests/test_models_gpu.py tests/test_kernels_gpu.py -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t16_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t16_tests.log
RPB_DX_SCATTER=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t16_bench_fused.log 2> gpurun_out/t16_bench_fused.err
RPB_DX_SCATTER=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t16_bench_sep.log 2> gpurun_out/t16_bench_sep.err
RPB_DX_SCATTER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t16_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t16_ncu_bench.log 2>&1
tail -8 gpurun_out/t16_tests.log | cut -c1-300; cut -c1-400 gpurun_out/t16_bench_fused.log; cut -c1-400 gpurun_out/t16_bench_sep.log; tail -3 gpurun_out/t16_bench_fused.err | cut -c1-300
