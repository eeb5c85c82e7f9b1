#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --no-cpu-baseline --no-train-step --no-extras 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fused ms',round(j['ms_per_step'],5), 'e2e', j['e2e']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:deepfm_fwd_fused -c 6 --csv --log-file gpurun_out/t42_fused.csv python bench.py --steps 2 --warmup 3 --eager --no-cpu-baseline --no-train-step --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/t42_fused.csv | tail -3 | cut -d, -f5,13- | cut -c1-200
( timeout 300 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k "one_kernel" ) 2>&1 | tail -2
