#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_linear_tc_gpu.py tests/test_models_gpu.py tests/test_kernels_gpu.py -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t18_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/t18_tests.log
RPB_DX_SCATTER=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t18_bench_fused.log 2> gpurun_out/t18_bench_fused.err
timeout 200 python tools/exp/trace_gemm.py > gpurun_out/t18_trace.log 2>&1
RPB_DX_SCATTER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t18_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t18_ncu_bench.log 2>&1
tail -8 gpurun_out/t18_tests.log | cut -c1-300; cut -c1-400 gpurun_out/t18_bench_fused.log; cat gpurun_out/t18_trace.log | head -30; tail -3 gpurun_out/t18_bench_fused.err | cut -c1-300
