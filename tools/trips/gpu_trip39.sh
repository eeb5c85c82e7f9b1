#!/bin/bash
mkdir -p gpurun_out
for rev in 1 0 1 0; do
RPB_OPTIONS=scatter_reverse=$rev timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('scatter_reverse',$rev,'ms',round(j['ms_per_step'],5))"
done
( timeout 300 python -m pytest tests/test_tower_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -p no:cacheprovider -k "deepfm or DeepFM" ) 2>&1 | tail -3
