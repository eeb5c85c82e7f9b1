#!/bin/bash
# round 2 (1 GPU): tcgen05 tower tail inside the 8-gather-warp one-kernel forward (default) — parity, trace, bench tc=1 / tc=0
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k 'one_kernel or fused_head' ) > gpurun_out/r2_03_fused_tests.log 2>&1
echo "fused tests exit $?" >> gpurun_out/r2_03_fused_tests.log
tail -12 gpurun_out/r2_03_fused_tests.log | cut -c1-600
timeout 300 python tools/exp/trace_fused.py > gpurun_out/r2_03_trace.log 2>&1; head -18 gpurun_out/r2_03_trace.log | cut -c1-200
for v in 1 0; do
  RPB_OPTIONS=fused_tc_tail=$v timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras --no-experiments 2> gpurun_out/r2_03_bench_tc$v.err \
    | tee gpurun_out/r2_03_bench_tc$v.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('tc_tail=$v ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'), 'frac', j['roofline'].get('frac'), 'loss', j['e2e']['loss'])"
done
( timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_03_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_03_tests.log
tail -8 gpurun_out/r2_03_tests.log | cut -c1-400
