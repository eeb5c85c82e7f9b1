#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_tower_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -p no:cacheprovider -k "one_kernel or DeepFM" ) 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --no-train-step --no-extras 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fused ms',round(j['ms_per_step'],5), 'e2e', j['e2e']['value'])"
timeout 200 python tools/exp/trace_fused.py 2>&1 | grep -v Warn | tail -26
