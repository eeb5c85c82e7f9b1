#!/bin/bash
mkdir -p gpurun_out/prof
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
for m in DeepFM xDeepFM AutoInt DCN FiBiNet MMOE; do
  timeout 400 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:rpb -c 400 --csv --log-file gpurun_out/prof/$m.csv python tools/profile_all.py --model $m --steps 2 > gpurun_out/prof/$m.log 2>&1
  echo "$m exit $?"; tail -1 gpurun_out/prof/$m.log | cut -c1-200
done
