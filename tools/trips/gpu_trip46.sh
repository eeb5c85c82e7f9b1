#!/bin/bash
# last call of round 1 (5.8 GPU-minutes left): (1) quick bench = validates the new in-step roofline leg of bench.py,
# (2) first hardware run of the tcgen05 tower tail (opt-in), (3) bench with it, (4) the full default bench if time remains.
mkdir -p gpurun_out
timeout 200 python bench.py --no-cpu-baseline --no-train-step --no-extras > gpurun_out/t46_quick.json 2> gpurun_out/t46_quick.err
echo "quick rc $?"; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/t46_quick.json').readline())
    r=j['roofline']
    print('ms/step', j['ms_per_step'], 'loss', j['e2e']['loss'], 'roofline kernel', r['kernel'][:40], 'us', r['us_per_launch'], 'frac', r['frac'], 'share', r.get('share_of_step'), 'err', r.get('fused_forward_error'))
    print('gather_only', {k:v for k,v in r.get('gather_only',{}).items() if k in ('us_per_launch','frac')})
except Exception as e:
    print('quick parse failed', e)
PY
grep -v Warning gpurun_out/t46_quick.err | tail -3 | cut -c1-300
( RPB_EXPERIMENTAL=1 timeout 100 python -m pytest tests/test_tower_gpu.py -m gpu -x -q -p no:cacheprovider -k tc_tail ) > gpurun_out/t46_tctail_test.log 2>&1
echo "tc_tail test rc $?"; tail -12 gpurun_out/t46_tctail_test.log | cut -c1-400
RPB_OPTIONS=fused_tc_tail=1 timeout 120 python bench.py --no-cpu-baseline --no-train-step --no-extras > gpurun_out/t46_tctail_bench.json 2> gpurun_out/t46_tctail_bench.err
echo "tc_tail bench rc $?"; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/t46_tctail_bench.json').readline())
    print('tc_tail ms/step', j['ms_per_step'], 'loss', j['e2e']['loss'], 'fwd us', j['roofline']['us_per_launch'])
except Exception as e:
    print('tc_tail bench parse failed', e)
PY
grep -v Warning gpurun_out/t46_tctail_bench.err | tail -3 | cut -c1-300
timeout 400 python bench.py > gpurun_out/t46_bench.json 2> gpurun_out/t46_bench.err
echo "full bench rc $?"; cut -c1-300 gpurun_out/t46_bench.json
