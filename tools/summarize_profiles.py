"""Summarise gpurun_out/prof/<model>.csv (ncu --csv metric dumps) into profiles/r01_kernels.md."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def load(path):
    rows = [l for l in open(path) if not l.startswith('==')]
    r = csv.reader(rows)
    hdr = next(r)
    ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    per = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        key = (row[ii], row[ki])
        try:
            per.setdefault(key, {})[row[mi]] = float(row[vi].replace(',', ''))
        except ValueError:
            pass
    return per


def main():
    out = ['# r01 — per-kernel ncu metrics at the BASELINE.json config shapes',
           '',
           'Command per model: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,'
           'dram__throughput...,sm__pipe_tensor_cycles_active...,sm__throughput... --clock-control none -k regex:rpb '
           'python tools/profile_all.py --model <M> --steps 2` (second step listed; times are ncu-serialised, cold cache).',
           f'HBM GB/s = (dram read + write) / duration; peak = {PEAK:.0f} GB/s measured copy bandwidth.', '']
    for m in ('DeepFM', 'xDeepFM', 'AutoInt', 'DCN', 'FiBiNet', 'MMOE'):
        p = os.path.join(ROOT, 'gpurun_out', 'prof', m + '.csv')
        if not os.path.exists(p):
            continue
        per = load(p)
        items = list(per.items())
        # keep the second half (second step)
        names = [k[1] for k, _ in items]
        half = len(items) // 2
        items = items[half:]
        agg = collections.OrderedDict()
        for (i, name), d in items:
            short = name.split('(')[0].replace('void ', '').replace('rpb::', '')
            a = agg.setdefault(short, {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0, 'tensor': 0.0, 'dram_pct': 0.0, 'regs': 0})
            a['n'] += 1
            a['us'] += d.get('gpu__time_duration.sum', 0) / 1e3
            a['rd'] += d.get('dram__bytes_read.sum', 0)
            a['wr'] += d.get('dram__bytes_write.sum', 0)
            a['tensor'] = max(a['tensor'], d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0))
            a['dram_pct'] = max(a['dram_pct'], d.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0))
            a['regs'] = int(d.get('launch__registers_per_thread', 0))
        tot = sum(a['us'] for a in agg.values())
        out += [f'## {m} (one training step: forward + backward + sparse grad re-zero; {tot:.0f} us of rpb kernels)', '',
                '| kernel | launches | time us | share | DRAM MB (rd+wr) | HBM GB/s | % of peak | dram thr % (max) | tensor pipe % (max) | regs |',
                '|---|---|---|---|---|---|---|---|---|---|']
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
            mb = (a['rd'] + a['wr']) / 1e6
            gbs = (a['rd'] + a['wr']) / (a['us'] * 1e-6) / 1e9 if a['us'] > 0 else 0
            out.append(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {mb:.1f} | {gbs:.0f} | "
                       f"{100 * gbs / PEAK:.1f}% | {a['dram_pct']:.1f} | {a['tensor']:.1f} | {a['regs']} |")
        out.append('')
    open(os.path.join(ROOT, 'profiles', 'r01_kernels.md'), 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:60]))


if __name__ == '__main__':
    main()
