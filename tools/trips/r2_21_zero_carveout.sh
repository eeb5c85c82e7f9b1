timeout 300 python tools/exp/overlap_zero.py 148 2>&1 | grep -v Warning | tail -4
timeout 600 python bench.py --zero-first 1 --no-cpu-baseline --no-train-step --no-extras 2>/dev/null | python -c "
import json,sys
j=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('zero_first=1 ms/step', round(j['ms_per_step'],4))"
