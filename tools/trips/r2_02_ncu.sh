#!/bin/bash
# round 2, call: microbenchmark of the row-fetch paths, then ncu launch list + full capture of the 8-warp fused forward
mkdir -p gpurun_out
timeout 200 ./tools/exp/exp_rowfetch > gpurun_out/r2_02_rowfetch.log 2>&1; cat gpurun_out/r2_02_rowfetch.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_02_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras --no-experiments > gpurun_out/r2_02_ncu_bench.log 2>&1
python tools/step_list.py gpurun_out/r2_02_launches.csv 2>&1 | tail -20
timeout 600 ncu --set full --import-source on --clock-control none -k regex:deepfm_fwd_fused --launch-skip 3 --launch-count 1 -o gpurun_out/r2_02_fused8 -f python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras --no-experiments > gpurun_out/r2_02_ncu_full.log 2>&1
tail -3 gpurun_out/r2_02_ncu_full.log
ncu -i gpurun_out/r2_02_fused8.ncu-rep --page raw --csv > gpurun_out/r2_02_fused8_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_02_fused8_raw.csv')))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else []
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_tensor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','l1tex__data_pipe_lsu_wavefronts','smsp__average_warps_issue_stalled','launch__registers_per_thread','sm__pipe_tensor']
for h,v in zip(hdr,vals):
    if any(w in h for w in want): print(h,'=',v)
PY
