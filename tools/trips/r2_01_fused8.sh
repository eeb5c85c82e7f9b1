#!/bin/bash
# round 2, call 1 (1 GPU): the 8-gather-warp one-kernel forward (deepfm_fwd_fused8_kernel, default) — parity vs the separate
# kernels, per-role trace, quick bench with 8 and with 4 gather warps, then the whole GPU suite.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k 'one_kernel or fused_head' ) > gpurun_out/r2_01_fused_tests.log 2>&1
echo "fused tests exit $?" >> gpurun_out/r2_01_fused_tests.log
tail -6 gpurun_out/r2_01_fused_tests.log | cut -c1-400
timeout 300 python tools/exp/trace_fused.py > gpurun_out/r2_01_trace8.log 2>&1; tail -30 gpurun_out/r2_01_trace8.log | cut -c1-200
for w in 8 4; do
  RPB_OPTIONS=fused_gather_warps=$w timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras --no-experiments 2> gpurun_out/r2_01_bench_w$w.err \
    | tee gpurun_out/r2_01_bench_w$w.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('gather_warps=$w ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'), 'frac', j['roofline'].get('frac'), 'loss', j['e2e']['loss'])"
done
for la in 3 4 5; do
  RPB_OPTIONS=fused_ring=$la timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras --no-experiments 2> gpurun_out/r2_01_bench_la$la.err \
    | tee gpurun_out/r2_01_bench_la$la.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('ring=$la ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'))"
done
( timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_01_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_01_tests.log
tail -8 gpurun_out/r2_01_tests.log | cut -c1-400
