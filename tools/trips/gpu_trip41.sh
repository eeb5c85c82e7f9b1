#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t41_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/t41_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:deepfm_fwd_fused --launch-skip 3 --launch-count 1 -o gpurun_out/t41_fused -f python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/t41_ncu_full.log 2>&1
tail -2 gpurun_out/t41_ncu_full.log
