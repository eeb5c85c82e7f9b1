#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/t25_tower.log 2>&1
echo "tower exit $?" >> gpurun_out/t25_tower.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/t25_bench.log 2> gpurun_out/t25_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t25_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t25_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tf32x3_v2 --launch-skip 6 --launch-count 1 -o gpurun_out/t25_fused -f python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t25_ncu_full.log 2>&1
RPB_TOWER_EPILOGUE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t25_launches_unfused.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline > gpurun_out/t25_ncu_bench2.log 2>&1
tail -4 gpurun_out/t25_tower.log | cut -c1-300; cut -c1-400 gpurun_out/t25_bench.log; tail -3 gpurun_out/t25_bench.err | cut -c1-300
