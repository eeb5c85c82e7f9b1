#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k "one_kernel" ) > gpurun_out/t40_fused.log 2>&1
echo "fused exit $?" >> gpurun_out/t40_fused.log
tail -30 gpurun_out/t40_fused.log | cut -c1-300
timeout 200 python bench.py --no-cpu-baseline --no-train-step --no-extras 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fused ms',round(j['ms_per_step'],5), 'e2e', j['e2e']['value'])"
RPB_FUSED_GATHER_GEMM=0 timeout 200 python bench.py --no-cpu-baseline --no-train-step --no-extras 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('unfused ms',round(j['ms_per_step'],5))"
