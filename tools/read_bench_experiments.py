#!/usr/bin/env python
"""Print the verdict of the `experiments` object of a bench.py JSON line (e.g. the driver's BENCH_rNN.json): which opt-in
variants passed parity on hardware and what they bought, i.e. which defaults to flip first.

    python tools/read_bench_experiments.py BENCH_r01.json
"""
import json
import sys


def find_line(obj):
    """The file may hold the line itself, a list of lines, or a wrapper dict with the line somewhere inside."""
    if isinstance(obj, dict):
        if 'experiments' in obj or 'ms_per_step' in obj:
            return obj
        for v in obj.values():
            r = find_line(v)
            if r is not None:
                return r
    if isinstance(obj, list):
        for v in obj:
            r = find_line(v)
            if r is not None:
                return r
    if isinstance(obj, str) and obj.lstrip().startswith('{'):
        try:
            return find_line(json.loads(obj))
        except Exception:
            return None
    return None


def main(path):
    text = open(path).read()
    try:
        line = find_line(json.loads(text))
    except Exception:
        line = None
        for ln in text.splitlines():
            if ln.lstrip().startswith('{'):
                line = find_line(json.loads(ln))
                if line is not None:
                    break
    if line is None:
        sys.exit('no bench line found in ' + path)
    print(f"headline: {line.get('ms_per_step')} ms/step, value {line.get('value')}, e2e {line.get('e2e', {}).get('value')}")
    r = line.get('roofline') or {}
    print(f"roofline kernel: {str(r.get('kernel'))[:60]}  us {r.get('us_per_launch')}  frac {r.get('frac')}  err {r.get('fused_forward_error')}")
    ex = line.get('experiments')
    if not ex:
        print('no experiments object (skipped or old bench)')
        return
    groups = {g: ex[g] for g in ('safe', 'tc_fwd', 'tc_bwd') if isinstance(ex.get(g), dict)} or {'all': ex}
    print(f"experiments: wall {ex.get('wall_s')} s in {len(groups)} child process(es)")
    for gname, grp in groups.items():
        report(gname, grp)


def report(gname, ex):
    base = ex.get('default_ms_per_step')
    print(f"[{gname}] wall {ex.get('wall_s')} s, default {base} ms/step, fwd {ex.get('default_fwd_us')} us, "
          f"infer fwd {ex.get('default_fwd_infer_us')} us; notes: {ex.get('timeout') or ex.get('skipped') or ex.get('error') or ex.get('default_error') or '-'}")
    for k, v in ex.items():
        if not isinstance(v, dict):
            continue
        ms = v.get('ms_per_step')
        gain = f'{100 * (base - ms) / base:+.1f} %' if (base and ms) else ''
        print(f"  {k:22s} parity_ok={v.get('parity_ok', '-')!s:5s} ms/step={ms} {gain}  "
              + ' '.join(f'{a}={b}' for a, b in v.items() if a not in ('ms_per_step', 'parity_ok', 'trace_cycles_cta0', 'tolerance', 'what')))
        if 'trace_cycles_cta0' in v:
            print('      trace:', v['trace_cycles_cta0'])


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'BENCH_r01.json')
