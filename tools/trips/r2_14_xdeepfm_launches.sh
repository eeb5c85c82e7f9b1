#!/bin/bash
# round 2: launch list of the xDeepFM step (config 3) with the CIN kernels on tcgen05, full 1-GPU suite, default bench line
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_14_xdeepfm_launches.csv \
    python bench.py --workload xdeepfm --steps 2 --warmup 1 --eager --no-cpu-baseline --no-train-step --no-extras > gpurun_out/r2_14_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = []
with open('gpurun_out/r2_14_xdeepfm_launches.csv') as f:
    lines = [l for l in f if l.startswith('"')]
r = list(csv.DictReader(lines))
per = collections.defaultdict(lambda: [0, 0.0])
for x in r:
    n = x['Kernel Name'][:70]
    try: v = float(x['Metric Value'].replace(',', ''))
    except Exception: continue
    per[n][0] += 1; per[n][1] += v
tot = sum(v[1] for v in per.values())
for n, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f'{t/1e3:10.1f} us  {c:4d}  {100*t/tot:5.1f}%  {n}')
print('total us', tot / 1e3, 'launches', len(r))
PY
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r2_14_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2_14_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_14_tests.log | tail -8 | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2_14_bench.json 2> gpurun_out/r2_14_bench.err
tail -c 1500 gpurun_out/r2_14_bench.json
