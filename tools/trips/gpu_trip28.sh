#!/bin/bash
mkdir -p gpurun_out
for st in 2 3; do for pw in 1 0; do
RPB_OPTIONS=wgrad_stages=$st RPB_PARALLEL_WGRAD=$pw timeout 300 python bench.py --no-cpu-baseline --no-train-step 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('stages',$st,'parallel',$pw,'ms',j['ms_per_step'])"
done; done
RPB_OPTIONS=wgrad_stages=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t28_launches.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step > gpurun_out/t28_ncu_bench.log 2>&1
