#!/bin/bash
# round 2: CIN second forward formulation (W.xk GEMM + x0 contraction), staged wgrad operands, hoisted de loads; A/B against cin_tc=2
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cin or xDeepFM or xdeepfm" ) > gpurun_out/r2_16_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_16_tests.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_16_tests.log | tail -12 | cut -c1-300
for tc in 1; do
  RPB_OPTIONS=cin_tc=$tc timeout 600 python bench.py --workload xdeepfm --steps 10 --warmup 3 > gpurun_out/r2_16_bench_tc$tc.json 2> gpurun_out/r2_16_bench_tc$tc.err
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open('gpurun_out/r2_16_bench_tc$tc.json') if l.startswith('{')][-1]
    print('cin_tc=$tc ms/step', round(j['ms_per_step'],3), 'e2e', j.get('e2e'), 'roofline', j.get('roofline'))
except Exception as e:
    print('no line', e)
PY
  tail -3 gpurun_out/r2_16_bench_tc$tc.err | cut -c1-300
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_16_xdeepfm_launches.csv \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_16_ncu_bench.log 2>&1
python tools/step_list.py gpurun_out/r2_16_xdeepfm_launches.csv 2>&1 | grep -i "cin\|total" | cut -c1-200
