#!/bin/bash
# round 2 (1 GPU): evidence for profiles/ — launch list of the bench step, ncu --set full of the forward kernel and of the two
# heaviest backward kernels, per-role trace, row-fetch microbenchmark, sanitizer summary
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_09_ncu_bench.log 2>&1
python tools/step_list.py gpurun_out/r02_bench_launches.csv 2>&1 | tail -16
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'deepfm_fwd_fs|gemm_tf32x3_v2|wgrad_tf32x3|tower_tail_bwd|rows_zero' --launch-skip 24 --launch-count 8 -o gpurun_out/r02_step -f python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_09_ncu_full.log 2>&1
tail -2 gpurun_out/r2_09_ncu_full.log
ncu -i gpurun_out/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/r02_step.ncu-rep 2>&1 | tail -14
timeout 300 python tools/exp/trace_fused.py > gpurun_out/r02_trace_fs.log 2>&1; head -26 gpurun_out/r02_trace_fs.log | cut -c1-220
timeout 200 ./tools/exp/exp_rowfetch > gpurun_out/r02_rowfetch.log 2>&1; cat gpurun_out/r02_rowfetch.log
bash tools/sanitize.sh 2>&1 | tail -5
