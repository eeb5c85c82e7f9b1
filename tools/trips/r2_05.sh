#!/bin/bash
# round 2 (1 GPU): fetch/split one-kernel forward (deepfm_fwd_fs_kernel): parity, trace, bench vs the 8-gather-warp kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k 'one_kernel or fused_head' ) > gpurun_out/r2_05_fused_tests.log 2>&1
echo "fused tests exit $?" >> gpurun_out/r2_05_fused_tests.log
tail -12 gpurun_out/r2_05_fused_tests.log | cut -c1-600
timeout 300 python tools/exp/trace_fused.py > gpurun_out/r2_05_trace.log 2>&1; head -27 gpurun_out/r2_05_trace.log | cut -c1-250
for v in 8 16 4; do
  RPB_OPTIONS=fused_fetch_warps=$v timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras 2> gpurun_out/r2_05_bench_f$v.err \
    | tee gpurun_out/r2_05_bench_f$v.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fetch_warps=$v ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'), 'frac', j['roofline'].get('frac'), 'loss', j['e2e']['loss'])"
  tail -3 gpurun_out/r2_05_bench_f$v.err | cut -c1-300
done
for la in 4; do
  RPB_OPTIONS=fused_ring=$la timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras 2>/dev/null \
    | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('ring=$la ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'))"
done
