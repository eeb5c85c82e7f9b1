#!/bin/bash
# NOT RUN YET (written at the end of round 1 with the GPU budget spent).  1 GPU:
#   gpurun --timeout 1500 -- 'bash tools/trips/r2_tc_tail.sh'
# Parity of the tcgen05 tower tail inside the one-kernel DeepFM forward (rpb_set_option fused_tc_tail=1) against the CUDA-core
# tail, then the bench with and without it and the per-role trace.  Every step runs under its own timeout: a protocol bug
# in the new mbarrier chain traps after ~2 s (mbar_wait is bounded) instead of hanging the box.
mkdir -p gpurun_out
RPB_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_tower_gpu.py tests/test_models_gpu.py -m gpu -q -p no:cacheprovider -k 'tc_tail or backward_tc or local_shards or autoint_vec or afm' 2>&1 | tail -15 | cut -c1-300
for v in 0 1; do
  RPB_OPTIONS=fused_tc_tail=$v,tower_bwd_tc=$v timeout 300 python bench.py --no-cpu-baseline --no-train-step --no-extras 2> gpurun_out/r2_tc_tail_$v.err \
    | tee gpurun_out/r2_tc_tail_$v.json | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('fused_tc_tail=$v ms/step', round(j['ms_per_step'],5), 'fwd us', j['roofline'].get('us_per_launch'), 'loss', j['e2e']['loss'])"
done
