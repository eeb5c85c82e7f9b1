#!/bin/bash
# round 2 (1 GPU): partial last round of the fused forward (per-CTA pieces), kink-exclusion parity tests, full suite, bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k 'one_kernel or fused_head' ) > gpurun_out/r2_04_fused_tests.log 2>&1
echo "fused tests exit $?" >> gpurun_out/r2_04_fused_tests.log
tail -12 gpurun_out/r2_04_fused_tests.log | cut -c1-600
timeout 300 python tools/exp/trace_fused.py > gpurun_out/r2_04_trace.log 2>&1; head -24 gpurun_out/r2_04_trace.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras --no-experiments 2> gpurun_out/r2_04_bench.err \
    | tee gpurun_out/r2_04_bench.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'), 'frac', j['roofline'].get('frac'), 'loss', j['e2e']['loss'])"
( timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_04_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_04_tests.log
grep -E "passed|failed|Error|FAILED" gpurun_out/r2_04_tests.log | tail -30 | cut -c1-400
