#!/bin/bash
# round-1 final profile refresh: bench JSON (both arms), launch list, ncu --set full of the DeepFM step, per-model metric tables
mkdir -p gpurun_out/prof
timeout 600 python bench.py > gpurun_out/t33_bench.json 2> gpurun_out/t33_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/t33_bench_ref.json 2> gpurun_out/t33_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t33_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step > gpurun_out/t33_ncu_bench.log 2>&1
# step 3 of the eager DeepFM loop: 9 hot kernels per step (gather, gemm+tail, tower bwd, 3 wgrad, dx+scatter, rows_zero) + 2 split_pack
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'wgrad_tf32x3|gemm_tf32x3_v2|gather_fwd_tile|rows_zero|tower_tail' --launch-skip 18 --launch-count 9 -o gpurun_out/t33_deepfm -f python tools/profile_all.py --model DeepFM --steps 3 > gpurun_out/t33_ncu_full.log 2>&1
tail -3 gpurun_out/t33_ncu_full.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
for m in DeepFM xDeepFM AutoInt DCN FiBiNet MMOE; do
  timeout 400 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:rpb -c 400 --csv --log-file gpurun_out/prof/$m.csv python tools/profile_all.py --model $m --steps 2 > gpurun_out/prof/$m.log 2>&1
  echo "$m exit $?"; tail -1 gpurun_out/prof/$m.log | cut -c1-200
done
cut -c1-600 gpurun_out/t33_bench.json; cut -c1-400 gpurun_out/t33_bench_ref.json
